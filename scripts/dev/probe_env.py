"""Quick device-time probe of the board kernels (not the bench contract; see bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.environment import BatchedGoEnv

def time_it(fn, reps=5):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

for n, Gs in ((9, (4096, 65536, 1 << 20)), (19, (512, 16384))):
    for G in Gs:
        for tile in (32, 8, 1):
            env = BatchedGoEnv(G, board_size=n, tile=tile)
            cnt = torch.zeros([1], dtype=torch.int64, device="cuda")
            def play():
                env.reset(); cnt.zero_(); env.playout(env.max_step + 1, seed=1, counter=cnt)
            ms = time_it(play, 3)
            plies = int(cnt.item())
            env.playout(40, seed=2)
            act = torch.full([G], n * n, dtype=torch.int16, device="cuda")
            ms_step = time_it(lambda: env.step(act))
            ms_enc = time_it(lambda: env.encode())
            ms_sc = time_it(lambda: env.score())
            print("n=%d G=%d tile=%d playout %.3f ms -> %.3e plies/s | step+legal %.3f ms (%.3e/s) encode %.3f ms (%.3e/s) score %.3f ms" % (
                n, G, tile, ms, plies / ms * 1e3, ms_step, G / ms_step * 1e3, ms_enc, G / ms_enc * 1e3, ms_sc), flush=True)
