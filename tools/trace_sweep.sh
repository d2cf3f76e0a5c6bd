cp die_e_b200/libdiee_cuda.so /tmp/keep.so; cp build_variants/libdiee_trace.so die_e_b200/libdiee_cuda.so
for cfg in "0 1" "0 4" "64 1" "64 2" "64 4" "80 2" "96 2"; do set -- $cfg; echo "== tree SMs $1 slices $2"; DIEE_TREE_SMS=$1 DIEE_SEARCH_SLICES=$2 python tools/trace_run.py; done
cp /tmp/keep.so die_e_b200/libdiee_cuda.so
