"""dev probe: stepwise kernels (step / legal / encode / score) per tile width and batch size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transgo_b200.environment import BatchedGoEnv

def best_ms(fn, reps=7):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

for n, Gs in ((9, (1024, 4096, 16384, 65536, 1 << 20)), (19, (512, 4096, 65536))):
    for G in Gs:
        row = []
        for tile in (32, 8, 1):
            if G * tile > (1 << 24):
                row.append("t%d: -" % tile); continue
            e = BatchedGoEnv(G, board_size=n, tile=tile)
            e.playout(60 if n == 9 else 200, seed=4)
            m = e.unpack_legal(e.legal_mask())
            acts = m.to(torch.uint8).argmax(1).to(torch.int16)
            st = e.states.clone()
            def step():
                e.states.copy_(st); e.step(acts)
            ms_copy = best_ms(lambda: e.states.copy_(st))
            ms_step = best_ms(step) - ms_copy
            ms_legal = best_ms(lambda: e.legal_mask())
            buf = torch.empty([G, 10, n, n], dtype=torch.float32, device="cuda")
            ms_enc = best_ms(lambda: e.encode(out=buf))
            ms_score = best_ms(lambda: e.score())
            row.append("t%d: step %.1f legal %.1f enc %.1f score %.1f us" % (tile, ms_step * 1e3, ms_legal * 1e3, ms_enc * 1e3, ms_score * 1e3))
            del e, buf
        print("n=%d G=%7d | %s" % (n, G, " | ".join(row)), flush=True)
