"""Developer probe: per-ply loop statistics of k_playout_rows.  Needs a stats build of the library:
   nvcc ... -DTG_ROWS_STATS -c transgo_b200/csrc/tg_env.cu ; link it with the other objects into gpurun_out/rows_stats.so"""
import sys, os, ctypes
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
_lib.LIB_PATH = "/root/repo/gpurun_out/rows_stats.so"
from transgo_b200.environment import BatchedGoEnv
L = _lib.lib()
names = ["warp-plies", "plies with enclosed pts", "T flood steps", "U iterations", "S flood steps", "capture floods entered", "capture flood steps", "capture not settled"]
for n, G in ((9, 4096), (19, 512)):
    env = BatchedGoEnv(G, board_size=n, tile=-1)
    out = (ctypes.c_ulonglong * 8)()
    L.tg_debug_rows_stats(out, 1)
    for lo, hi in ((0, 40), (40, 80), (80, 200)) if n == 9 else ((0, 200), (200, 500), (500, 800)):
        env.playout(hi - lo, seed=1)
        L.tg_debug_rows_stats(out, 1)
        v = list(out)
        print("n=%d plies %d-%d:" % (n, lo, hi), "; ".join("%s %.2f" % (nm, x / max(1, v[0])) for nm, x in zip(names, v)), "(per warp-ply; warp-plies=%d)" % v[0])
