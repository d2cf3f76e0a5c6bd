# sweep of the SM partition of the sliced search (tree SMs x slices), BASELINE configs[2] on one GPU
run() { python bench.py --no-subrecords --no-cpu-baseline --no-parity-check --steps 10 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  ', d['value'], d['ms_per_step'], d['e2e']['value'])"; }
echo "no partition (DIEE_TREE_SMS=0)"; DIEE_TREE_SMS=0 run
for t in 48 64 80 96; do for s in 2 3 4; do echo "tree SMs $t slices $s"; DIEE_TREE_SMS=$t DIEE_SEARCH_SLICES=$s run; done; done
