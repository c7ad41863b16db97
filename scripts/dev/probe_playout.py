"""Tile-width sweep of the playout kernel (device time, CUDA events; not the bench contract)."""
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


ref = {}
for n, Gs in ((9, (4096, 16384, 65536, 1 << 20)), (19, (512, 8192))):
    for G in Gs:
        for tile in (-1, 32, 16, 16 | 64, 8, 8 | 64, 4 | 64, 1):
            if G >= 65536 and tile > 8:
                continue
            env = BatchedGoEnv(G, board_size=n, tile=tile)
            cnt = torch.zeros([1], dtype=torch.int64, device="cuda")

            def play():
                env.reset(); cnt.zero_(); env.playout(env.max_step + 1, seed=1, counter=cnt)
            ms = time_it(play, 4)
            h = int(env.states.to(torch.int64).sum().item())
            ref.setdefault((n, G), h)
            print("n=%d G=%d tile=%d playout %.3f ms -> %.3e plies/s  state-hash %s" % (
                n, G, tile, ms, int(cnt.item()) / ms * 1e3, "same" if ref[(n, G)] == h else "DIFFERENT"), flush=True)
            del env
