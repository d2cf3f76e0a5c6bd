# 8-GPU evidence run: the default line (what the driver's scaling step runs), C5 (16,384 games per GPU x 8 = 131,072
# concurrent self-play games, time-boxed, trajectories all-gathered through the C ABI), playouts
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_mcts_n8.json 2> gpurun_out/r02_bench_mcts_n8.err
timeout 300 $TR bench.py --gpus 8 --workload selfplay --games 16384 --precision bf16 --max-waves 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_c5_bf16_n8.json 2> gpurun_out/r02_bench_c5_bf16_n8.err
timeout 400 $TR bench.py --gpus 8 --workload selfplay --games 16384 --precision split3 --max-waves 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_c5_split3_n8.json 2> gpurun_out/r02_bench_c5_split3_n8.err
timeout 200 $TR bench.py --gpus 8 --workload playout > gpurun_out/r02_bench_playout_n8.json 2> gpurun_out/r02_bench_playout_n8.err
for f in mcts c5_bf16 c5_split3 playout; do grep -E "Error|FAILED|error" gpurun_out/r02_bench_${f}_n8.err | head -3; grep "^{" gpurun_out/r02_bench_${f}_n8.json | cut -c1-260; done
