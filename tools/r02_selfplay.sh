# whole self-play runs (reference mode: every game to its end) per precision, the non-parity modes, and the large batches
set -x
python bench.py --workload selfplay --precision bf16 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_bench_selfplay_bf16.json 2> gpurun_out/r02_sp1.err
python bench.py --workload selfplay --precision split3 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_bench_selfplay_split3.json 2> gpurun_out/r02_sp2.err
python bench.py --workload selfplay --precision bf16 --refill 3072 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_bench_selfplay_bf16_refill.json 2> gpurun_out/r02_sp3.err
python bench.py --workload selfplay --precision bf16 --refill 3072 --leaves 4 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_bench_selfplay_bf16_refill_vl4.json 2> gpurun_out/r02_sp4.err
python bench.py --workload selfplay --precision split3 --refill 1200 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_bench_selfplay_split3_refill.json 2> gpurun_out/r02_sp5.err
python bench.py --workload alpha --precision bf16 --leaves 4 --no-cpu-baseline > gpurun_out/r02_bench_alpha_bf16_vl4.json 2> gpurun_out/r02_sp6.err
for g in 65536; do python bench.py --no-subrecords --no-cpu-baseline --no-parity-check --steps 3 --warmup 2 --games $g | cut -c1-200; done
for f in selfplay_bf16 selfplay_split3 selfplay_bf16_refill selfplay_bf16_refill_vl4 selfplay_split3_refill alpha_bf16_vl4; do python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/r02_bench_$f.json') if l.startswith('{')][0]); dd=d['detail']; print('$f', d['value'], d['unit'], d['ms_per_step'], dd.get('game_moves_per_sec'), dd.get('simulations_per_sec'), d['roofline']['frac'], dd.get('waves_per_step'), dd.get('games_finished_per_step'))"; done
