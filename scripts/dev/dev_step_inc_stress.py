"""dev: long randomized cross-check of tg_env_step_inc against tg_env_step (illegal moves, both PASS spellings, resign, restarts
of finished games, stale / garbage atari records), several seeds and both board sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.environment import BatchedGoEnv
total = 0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    for n, G, plies in ((9, 3000, 260), (19, 200, 500)):
        gen = torch.Generator(device="cuda"); gen.manual_seed(seed * 17 + n)
        a = BatchedGoEnv(G, board_size=n, step_engine="incremental"); b = BatchedGoEnv(G, board_size=n, step_engine="analysis")
        P = n * n
        for ply in range(plies):
            legal = b.unpack_legal(b.legal_mask())
            acts = (torch.rand([G, P + 1], device="cuda", generator=gen) * legal).argmax(1).to(torch.int16)
            r = torch.rand([G], device="cuda", generator=gen)
            acts = torch.where(r < 0.05, torch.randint(0, P, [G], device="cuda", generator=gen).to(torch.int16), acts)      # maybe illegal
            acts = torch.where((r >= 0.05) & (r < 0.08), torch.full_like(acts, P), acts)                                       # PASS as P
            acts = torch.where((r >= 0.08) & (r < 0.10), torch.full_like(acts, -1), acts)                                      # PASS as -1
            acts = torch.where((r >= 0.10) & (r < 0.101), torch.full_like(acts, -2), acts)                                     # resign
            acts = torch.where((r >= 0.101) & (r < 0.103), torch.full_like(acts, P + 5), acts)                                 # out of range
            la, da = a.step(acts); lb, db = b.step(acts)
            if not (torch.equal(a.states, b.states) and torch.equal(la, lb) and torch.equal(da, db) and torch.equal(a.accepted, b.accepted)):
                print("MISMATCH seed", seed, "n", n, "ply", ply); sys.exit(1)
            total += G
            if ply % 60 == 59:                                       # restart the finished games from scratch on both sides
                fin = da.bool()
                fresh = BatchedGoEnv(G, board_size=n).states
                a.states[:, fin] = fresh[:, fin]; b.states[:, fin] = fresh[:, fin]
            if ply % 97 == 50:
                a.atari.random_(0, 2 ** 31 - 1)                      # garbage records must be noticed
print("OK: %d game-steps identical" % total)
