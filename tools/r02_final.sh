# final evidence run of round 2 (on the GPU box) -> gpurun_out/*_k.*
T=k
python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_$T.txt 2>&1; tail -2 gpurun_out/r02_pytest_gpu_$T.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$T.txt 2>&1; tail -1 gpurun_out/r02_smoke_$T.txt
python bench.py > gpurun_out/r02_bench_mcts_$T.json 2> gpurun_out/r02_bench_mcts_$T.err; tail -c 300 gpurun_out/r02_bench_mcts_$T.err
python bench.py --impl reference > gpurun_out/r02_bench_ref_$T.json 2> gpurun_out/r02_bench_ref_$T.err
python bench.py --workload playout > gpurun_out/r02_bench_playout_$T.json 2> gpurun_out/r02_bench_playout_$T.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_mcts_$T.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lane_pack_kernel -s 1 -c 1 -o gpurun_out/r02_prof_packhead_$T -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-subrecords --no-parity-check > gpurun_out/ncu_f1.log 2>&1
ncu -i gpurun_out/r02_prof_packhead_$T.ncu-rep --page raw --csv > gpurun_out/r02_prof_packhead_${T}_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:lane_pack_kernel -s 1 -c 1 -o gpurun_out/r02_prof_packplay_$T -f python bench.py --workload playout --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check > gpurun_out/ncu_f2.log 2>&1
ncu -i gpurun_out/r02_prof_packplay_$T.ncu-rep --page raw --csv > gpurun_out/r02_prof_packplay_${T}_raw.csv 2>/dev/null
ls -la gpurun_out | grep "_$T"
