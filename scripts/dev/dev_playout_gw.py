"""dev: incremental playout, games per warp below the full warp (empty upper tiles): launch time and result check."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
from transgo_b200.environment import BatchedGoEnv
L = _lib.lib()
L.tg_env_debug_games_per_warp.argtypes = [ctypes.c_int]


def t(env, reps=6):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.playout(env.max_step + 1, seed=1, from_reset=True); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for n, G in ((9, 4096), (9, 2048), (9, 8192), (19, 512)):
    ref = None
    for lanes in ((4, 8) if n == 9 else (32, 16, 8, 4)):
        for gw in [0] + [g for g in (6, 5, 4, 3, 2, 1) if g < 32 // lanes]:
            L.tg_env_debug_games_per_warp(gw)
            env = BatchedGoEnv(G, board_size=n, tile=lanes | 128)
            ms = t(env)
            if ref is None:
                ref = env.states.clone()
            same = bool(torch.equal(env.states, ref))
            print("n=%d G=%d lanes=%d games/warp=%s: %.4f ms  %.3e plies/s  identical=%s" % (n, G, lanes, gw or "full", ms, G * (120 if n == 9 else 722) / ms * 1e3, same), flush=True)
L.tg_env_debug_games_per_warp(0)
