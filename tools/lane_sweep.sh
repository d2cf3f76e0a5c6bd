# sweep of the lane kernel's residency / batching knobs on the C3 rollouts and C2 playouts (run on the GPU box)
for bps in 2 3 4 5 6 8 12; do for sm in 8 16 24; do
  export DIEE_LANE_FORCE_BPS=$bps DIEE_LANE_STORE_MIN=$sm
  echo "bps=$bps store_min=$sm"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-large-batch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' mcts', d['value'], d['config']['rollout_kernel_ms'])"
  python bench.py --workload playout --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' playout', d['value'], d['ms_per_step'])"
done; done
