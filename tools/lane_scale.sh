# resident CTAs per SM of the rollout kernel around the default (run on the GPU box)
for b in 9 10 11 12; do
DIEE_LANE_BLOCKS_PER_SM=$b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-large-batch 2>gpurun_out/slice_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bps', $b, 'mcts', d['value'], d['ms_per_step'], d['config']['tree_kernel_ms'], d['config']['rollout_kernel_ms'])"
done
