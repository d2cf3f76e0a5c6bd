# usage: bash tools/variant_run.sh <lib.so> <bench args...>   -- runs bench.py with another build of the library
LIB=$1; shift
cp die_e_b200/libdiee_cuda.so /tmp/libdiee_orig.so
cp $LIB die_e_b200/libdiee_cuda.so
touch die_e_b200/libdiee_cuda.so
python bench.py "$@"
cp /tmp/libdiee_orig.so die_e_b200/libdiee_cuda.so
touch die_e_b200/libdiee_cuda.so
