# round-2 final evidence run (on the GPU box) -> gpurun_out/*_g.*   usage: bash tools/r02_capture_g.sh
TAG=${1:-g}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_$TAG.txt 2>&1; tail -2 gpurun_out/r02_pytest_gpu_$TAG.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$TAG.txt 2>&1; tail -1 gpurun_out/r02_smoke_$TAG.txt
( time python bench.py > gpurun_out/r02_bench_mcts_$TAG.json 2> gpurun_out/r02_bench_mcts_$TAG.err ) 2>&1 | grep real
python bench.py --impl reference > gpurun_out/r02_bench_ref_$TAG.json 2> gpurun_out/r02_bench_ref_$TAG.err
python bench.py --workload playout > gpurun_out/r02_bench_playout_$TAG.json 2> gpurun_out/r02_bench_playout_$TAG.err


# launch lists: the headline step, and the default line with its sub-records (the large batches run lane_pack_kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_mcts_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_l1.log 2>&1
# full captures: the packed rollout kernel on 8,192 games, the lane-resident one and the tree kernel on the headline batch
ncu --set full --clock-control none --import-source on -k regex:lane_pack_kernel -s 1 -c 1 -o gpurun_out/r02_prof_pack_$TAG -f python tools/pack_stats.py 8192 > gpurun_out/ncu_f0.log 2>&1
ncu -i gpurun_out/r02_prof_pack_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_prof_pack_${TAG}_raw.csv 2>/dev/null
grep "^games" gpurun_out/ncu_f0.log
ncu --set full --clock-control none --import-source on -k regex:lane_run_kernel -s 9 -c 1 -o gpurun_out/r02_prof_lane_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_f1.log 2>&1
ncu -i gpurun_out/r02_prof_lane_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_prof_lane_${TAG}_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:mcts_search_kernel -s 1 -c 1 -o gpurun_out/r02_prof_tree_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_f2.log 2>&1
ncu -i gpurun_out/r02_prof_tree_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_prof_tree_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out | grep "_$TAG" 
