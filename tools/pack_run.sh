#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_lane_pack.py tests/test_gpu_baseline_sizes.py -q -x -k "not c4 and not self_play" 2>&1 | tail -2
for n in 512 768 1024 1536 2048 4096 8192; do echo -n "default "; timeout 60 python tools/pack_stats.py $n | head -1; done
echo -n "lane-resident "; DIEE_LANE_PACK=0 timeout 60 python tools/pack_stats.py 768 | head -1
echo -n "lane-resident "; DIEE_LANE_PACK=0 timeout 60 python tools/pack_stats.py 1536 | head -1
for p in 0 2; do DIEE_LANE_PACK=$p python bench.py --workload playout --steps 10 --warmup 3 --no-subrecords 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('playout PACK=$p', r['value'], r['ms_per_step'], r.get('parity_check', {}).get('ok'))"; done
python bench.py --no-subrecords 2>/dev/null | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('headline', r['value'], r['ms_per_step'], r['e2e']['value'], r['parity_check']['ok'], r['roofline']['kernel'], r['roofline']['frac'])"
