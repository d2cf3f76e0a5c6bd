# round-end evidence run (on the GPU box): tests, smoke, bench lines, ncu launch lists and full captures -> gpurun_out/
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
python bench.py > gpurun_out/bench_mcts.json 2> gpurun_out/bench_mcts.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --workload playout > gpurun_out/bench_playout.json 2> gpurun_out/bench_playout.err
python bench.py --rollout check_current --no-large-batch --no-cpu-baseline > gpurun_out/bench_mcts_cc.json 2> gpurun_out/bench_cc.err
python bench.py --workload alpha --no-cpu-baseline > gpurun_out/bench_alpha.json 2> gpurun_out/bench_alpha.err
python bench.py --workload selfplay --no-cpu-baseline > gpurun_out/bench_selfplay.json 2> gpurun_out/bench_selfplay.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_mcts.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-large-batch > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_playout.csv python bench.py --workload playout --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lane_run_kernel -s 9 -c 1 -o gpurun_out/prof_lane_r01i -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-large-batch > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mcts_search_kernel -s 1 -c 1 -o gpurun_out/prof_tree_r01i -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-large-batch > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lane_run_kernel -s 10 -c 1 -o gpurun_out/prof_playout_r01i -f python bench.py --workload playout --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out | tail -30
