"""dev: tg_env_step_inc vs tg_env_step: agreement on random games and time per step at several batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.environment import BatchedGoEnv
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
for n, G in ((9, 4096), (9, 65536), (9, 1 << 20), (19, 512), (19, 16384)):
    a = BatchedGoEnv(G, board_size=n, step_engine="incremental"); b = BatchedGoEnv(G, board_size=n, step_engine="analysis")
    ok = True
    ta = tb = 0.0
    plies = 60
    for ply in range(plies):
        legal = b.unpack_legal(b.legal_mask())
        acts = (torch.rand([G, n * n + 1], device="cuda", generator=gen) * legal).argmax(1).to(torch.int16)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        ev[0].record(); la, da = a.step(acts); ev[1].record(); lb, db = b.step(acts); ev[2].record(); torch.cuda.synchronize()
        ta += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
        ok = ok and bool(torch.equal(a.states, b.states) and torch.equal(la, lb) and torch.equal(da, db))
    print("n=%d G=%d: identical=%s  incremental %.1f us/step (%.3e steps/s)  analysis %.1f us/step (%.3e steps/s)" % (
        n, G, ok, ta / plies * 1e3, G / (ta / plies) * 1e3, tb / plies * 1e3, G / (tb / plies) * 1e3), flush=True)
