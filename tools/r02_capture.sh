# round-2 evidence run (on the GPU box): tests, smoke, bench lines, ncu launch list and full captures -> gpurun_out/
# usage: bash tools/r02_capture.sh [tag]
TAG=${1:-a}
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_$TAG.txt 2>&1; tail -5 gpurun_out/r02_pytest_gpu_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$TAG.txt 2>&1; tail -2 gpurun_out/r02_smoke_$TAG.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_mcts_$TAG.json 2> gpurun_out/r02_bench_mcts_$TAG.err; tail -c 600 gpurun_out/r02_bench_mcts_$TAG.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_$TAG.json 2> gpurun_out/r02_bench_ref_$TAG.err
python bench.py --workload playout > gpurun_out/r02_bench_playout_$TAG.json 2> gpurun_out/r02_bench_playout_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_mcts_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lane_run_kernel -s 9 -c 1 -o gpurun_out/r02_prof_lane_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_f1.log 2>&1
ncu -i gpurun_out/r02_prof_lane_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_prof_lane_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
