import ctypes, numpy as np, sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import positions, orc
from die_e_b200 import _ffi
ctx=_ffi.Context(0)
lib=_ffi.lib()
out=(ctypes.c_ulonglong*16)()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
states = positions.midgame_positions(seed=0xD1EE, n=N, max_adv=80)
cfg = orc.mcts_cfg(iterations=100, c=2.0, limit=400, mode=_ffi.MODE_PASS_CHILD)
lib.diee_debug_lane_stats(out,1)
ctx.mcts_search(_ffi.GAME_BACKGAMMON, states, states["player"].copy(), cfg, 0xD1EE, 0, 0)
lib.diee_debug_lane_stats(out,1)
names=['done','closed','walk','store']
for p in range(1,4):
    print(names[p], 'steps', out[p], 'lanes', out[8+p], 'avg lanes/step %.1f'%(out[8+p]/max(out[p],1)))
v = max(out[4], 1)
print('votes', out[4], 'per vote: idle %.1f closed %.1f walk %.1f store %.1f' % (out[5] / v, out[6] / v, out[7] / v, out[12] / v))
print('ply executions (warp level)', out[13], 'lanes %.1f' % (out[14] / max(out[13], 1)))
