"""Playout launch time vs batch size for each layout: flat = bound by the 120-ply dependency chain, linear = by throughput."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.environment import BatchedGoEnv


def t(env, reps=4):
    best = 1e9
    for _ in range(reps):
        env.reset(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.playout(env.max_step + 1, seed=1); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


print("games   " + "  ".join("%8s" % x for x in ("T=32", "T=16", "WU16", "WU8", "WU4", "rows")))
for G in (148, 592, 1184, 2368, 4096, 8192, 16384):
    row = []
    for tile in (32, 16, 16 | 64, 8 | 64, 4 | 64, -1):
        row.append(t(BatchedGoEnv(G, tile=tile)))
    print("%6d  " % G + "  ".join("%6.3fms" % x for x in row), flush=True)
