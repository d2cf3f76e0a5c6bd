# round-2 evidence run (on the GPU box): tests, smoke, bench lines, ncu launch lists and full captures -> gpurun_out/
# usage: bash tools/r02_capture.sh [tag]
TAG=${1:-a}
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -q -m gpu -s > gpurun_out/r02_pytest_gpu_$TAG.txt 2>&1; tail -3 gpurun_out/r02_pytest_gpu_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$TAG.txt 2>&1; tail -2 gpurun_out/r02_smoke_$TAG.txt
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_mcts_$TAG.json 2> gpurun_out/r02_bench_mcts_$TAG.err ) 2>&1 | grep real
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_$TAG.json 2> gpurun_out/r02_bench_ref_$TAG.err
python bench.py --workload playout > gpurun_out/r02_bench_playout_$TAG.json 2> gpurun_out/r02_bench_playout_$TAG.err
python bench.py --workload alpha --precision split3 --no-cpu-baseline > gpurun_out/r02_bench_alpha_split3_$TAG.json 2> gpurun_out/r02_bench_alpha_$TAG.err
python bench.py --workload alpha --precision bf16 --no-cpu-baseline > gpurun_out/r02_bench_alpha_bf16_$TAG.json 2>> gpurun_out/r02_bench_alpha_$TAG.err
python bench.py --workload alpha --precision fp32 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r02_bench_alpha_fp32_$TAG.json 2>> gpurun_out/r02_bench_alpha_$TAG.err
# launch lists (gpu__time_duration only) of the default line and of one AlphaZero search
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_mcts_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r02_launches_alpha_split3_$TAG.csv python bench.py --workload alpha --precision split3 --steps 1 --warmup 0 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_l2.log 2>&1
# full captures: the dominant kernel of the default line, the tree kernel, one tower convolution of each precision
ncu --set full --clock-control none --import-source on -k regex:lane_run_kernel -s 9 -c 1 -o gpurun_out/r02_prof_lane_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_f1.log 2>&1
ncu -i gpurun_out/r02_prof_lane_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_prof_lane_${TAG}_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:mcts_search_kernel -s 1 -c 1 -o gpurun_out/r02_prof_tree_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_f2.log 2>&1
ncu -i gpurun_out/r02_prof_tree_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_prof_tree_${TAG}_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 40 -c 2 -o gpurun_out/r02_prof_conv_split3_$TAG -f python bench.py --workload alpha --precision split3 --steps 1 --warmup 0 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_f3.log 2>&1
ncu -i gpurun_out/r02_prof_conv_split3_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_prof_conv_split3_${TAG}_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 45 -c 1 -o gpurun_out/r02_prof_conv_bf16_$TAG -f python bench.py --workload alpha --precision bf16 --steps 1 --warmup 0 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_f4.log 2>&1
ncu -i gpurun_out/r02_prof_conv_bf16_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_prof_conv_bf16_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | tail -25
python tools/net_lat.py > gpurun_out/r02_net_latency_$TAG.txt 2>&1; cat gpurun_out/r02_net_latency_$TAG.txt
