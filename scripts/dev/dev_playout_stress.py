"""dev: randomized agreement check of the playout layouts (library choice vs the row-per-lane formulation vs one thread per game
vs 8 / 32 lanes), random batch sizes, seeds, no-eye rule, ply counts, both board sizes; from reset and from mid-game positions."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.environment import BatchedGoEnv
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
total = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    n = 9 if random.random() < 0.75 else 19
    G = random.choice([1, 7, 8, 9, 63, 500, 4096, 4097, random.randint(1, 30000)]) if n == 9 else random.choice([1, 5, 64, 300, random.randint(1, 1500)])
    seed, no_eye = random.getrandbits(48), random.random() < 0.4
    plies1, plies2 = random.randint(0, 60), random.randint(1, (120 if n == 9 else 722) + 1)
    ref = None
    for tile in (0, -1, 1 | 128, 8 | 128, 32 | 128, 4 | 64):
        if n == 19 and tile == (4 | 64) and G > 800:
            continue
        env = BatchedGoEnv(G, board_size=n, tile=tile)
        tr = torch.full([plies1 + plies2, G], -9, dtype=torch.int16, device="cuda")
        sc = torch.empty([G], dtype=torch.float32, device="cuda")
        env.playout(plies1, seed=seed, no_eye=no_eye, trace=tr[:plies1]) if plies1 else None       # a mid-game position first
        env.playout(plies2, seed=seed + 1, no_eye=no_eye, trace=tr[plies1:], score_out=sc)          # (the incremental engine rebuilds its atari set)
        cur = (env.states.clone(), tr, sc)
        if ref is None:
            ref = cur
        elif not all(torch.equal(a, b) for a, b in zip(ref, cur)):
            print("MISMATCH n=%d G=%d seed=%d no_eye=%s plies=%d+%d tile=%d" % (n, G, seed, no_eye, plies1, plies2, tile)); sys.exit(1)
    total += G * (plies1 + plies2)
print("OK: %d game-plies, six layouts identical" % total)
