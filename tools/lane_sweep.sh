for bps in 3 4 6; do for lag in 0 1; do
  export DIEE_LANE_LAG=$lag DIEE_LANE_BLOCKS_PER_SM=$bps
  echo "bps=$bps lag=$lag"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-130
  python bench.py --workload playout --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-130
done; done
