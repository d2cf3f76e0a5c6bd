# round-end sanity run (on the GPU box): the driver's own sequence -- GPU tests, smoke, the two bench arms -- plus the playout line
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; tail -2 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_mcts.json 2> gpurun_out/bench_mcts.err
python bench.py --workload playout > gpurun_out/bench_playout.json 2> gpurun_out/bench_playout.err
python bench.py --workload alpha --no-cpu-baseline > gpurun_out/bench_alpha.json 2> gpurun_out/bench_alpha.err
tail -c 300 gpurun_out/bench_mcts.err
