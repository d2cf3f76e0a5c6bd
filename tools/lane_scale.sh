# store/refill batching at a refilled job size (run on the GPU box)
for sm in 4 8 16 24 32; do
DIEE_LANE_STORE_MIN=$sm python bench.py --games 8192 --steps 3 --warmup 2 --no-cpu-baseline --no-large-batch 2>gpurun_out/slice_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('store_min', $sm, 'mcts', d['value'], d['ms_per_step'], d['config']['tree_kernel_ms'], d['config']['rollout_kernel_ms'])"
done
