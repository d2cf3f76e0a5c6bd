import sys, numpy as np, argparse
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import orc, bench
from die_e_b200 import _ffi as ffi
orc.build()
ctx=ffi.Context(0)
s=bench.initial_states(ffi,0,64)
want=bench.host_states_for_reference(argparse.Namespace(games=64,workload="mcts"),0,64)
for rep in range(3):
    got=bench.midgame_states(ctx,ffi,0,64)
    print("rep",rep,"midgame diff games", [g for g in range(64) if got[g:g+1].tobytes()!=want[g:g+1].tobytes()])
g=7
print("got ",got[g]); print("want",want[g])
w,p,f=ctx.bg_playout(s,seed=bench.SEED,first_game_id=0,round_limit=70,want_finals=True)
print("gpu70",f[g], p[g], w[g])
ow,op,of=orc.bg_playout(s[g:g+1],bench.SEED,g,70)
print("orc70",of[0], op, ow)
a=s[g:g+1].copy()
for ply in range(70):
    if orc.bg_check_winner(a) is not None: print("winner at", ply); break
    orc.bg_random_ply(a, orc.philox(bench.SEED,ply,g,orc.STREAM_GAME,0))
print("loop70",a[0])
