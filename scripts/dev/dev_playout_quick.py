"""dev: headline playout launch time (4096 games, library layout) + the all-layouts-agree check at a few sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.environment import BatchedGoEnv


def t(env, reps=8):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.playout(env.max_step + 1, seed=1, from_reset=True); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for n, G in ((9, 4096), (9, 2048), (9, 16384), (9, 65536), (19, 512), (19, 4096)):
    env = BatchedGoEnv(G, board_size=n)
    ms = t(env)
    ref = BatchedGoEnv(G, board_size=n, tile=-1 if n == 9 else 32)          # an independent formulation (rows / bitboard tiles)
    ref.reset(); ref.playout(ref.max_step + 1, seed=1)
    print("n=%d G=%d %s: %.4f ms  %.3e plies/s  equal to the %s layout: %s" % (n, G, env.playout_layout(), ms, G * (120 if n == 9 else 722) / ms * 1e3,
          "rows" if n == 9 else "32-lane bitboard", bool(torch.equal(env.states, ref.states))), flush=True)
