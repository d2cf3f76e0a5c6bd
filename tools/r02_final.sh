python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_j.txt 2>&1; tail -2 gpurun_out/r02_pytest_gpu_j.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_j.txt 2>&1; tail -1 gpurun_out/r02_smoke_j.txt
python bench.py > gpurun_out/r02_bench_mcts_j.json 2> gpurun_out/r02_bench_mcts_j.err; tail -c 300 gpurun_out/r02_bench_mcts_j.err
python bench.py --impl reference > gpurun_out/r02_bench_ref_j.json 2> gpurun_out/r02_bench_ref_j.err
python bench.py --workload playout > gpurun_out/r02_bench_playout_j.json 2> gpurun_out/r02_bench_playout_j.err
