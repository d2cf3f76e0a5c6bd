# how the search scales with the job (run on the GPU box): games per launch, ply cap, slices on side streams,
# store/refill batch at a refilled job size.  Prints tree / rollout kernel times from the library's own events.
P='import sys,json; d=json.loads(sys.stdin.read()); c=d["config"]; print(sys.argv[1], d["value"], d["ms_per_step"], c.get("tree_kernel_ms"), c.get("rollout_kernel_ms"), c.get("rollout_plies_per_simulation"))'
for g in 128 256 512 1024 2048 4096; do
  python bench.py --games $g --steps 5 --warmup 3 --no-cpu-baseline --no-large-batch 2>/dev/null | python -c "$P" "games=$g"
done
for l in 100 200 300; do
  python bench.py --round-limit $l --steps 5 --warmup 3 --no-cpu-baseline --no-large-batch 2>/dev/null | python -c "$P" "limit=$l"
done
for s in 1 2 4; do
  DIEE_SEARCH_SLICES=$s python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-large-batch 2>/dev/null | python -c "$P" "slices=$s"
done
for sm in 4 8 16 24 32; do
  DIEE_LANE_STORE_MIN=$sm python bench.py --games 8192 --steps 3 --warmup 2 --no-cpu-baseline --no-large-batch 2>/dev/null | python -c "$P" "store_min=$sm"
done
