"""dev probe: the incremental playout engine (tile | 128) -- parity on every game, then a timing table."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from transgo_b200.environment import BatchedGoEnv
from oracle.go_oracle import OracleEnv

def parity(n, G, tile, no_eye=False, seed=0x7261):
    env = BatchedGoEnv(G, board_size=n, tile=tile)
    plies = env.max_step + 2
    trace = torch.full([plies, G], -9, dtype=torch.int16, device="cuda")
    cnt = torch.zeros([1], dtype=torch.int64, device="cuda")
    env.playout(plies, seed=seed, no_eye=no_eye, trace=trace, counter=cnt)
    torch.cuda.synchronize()
    o = OracleEnv(n, max_step=env.max_step)
    try:
        nn, finals = o.replay_playout(trace.cpu().numpy(), seed, 0, no_eye)
    except AssertionError as e:
        print("PARITY FAIL n=%d G=%d tile=%d no_eye=%d: %s" % (n, G, tile, no_eye, e)); return False
    ok = nn == int(cnt.item())
    planes, score, terr, legal, noeye = o.batch_observe(finals)
    ok = ok and np.array_equal(env.score().cpu().numpy(), score) and np.array_equal(env.encode().cpu().numpy(), planes)
    ref = BatchedGoEnv(G, board_size=n, tile=32); ref.playout(plies, seed=seed, no_eye=no_eye)
    ok = ok and torch.equal(ref.states, env.states)
    # resumability: the atari set is rebuilt from scratch at every launch
    two = BatchedGoEnv(G, board_size=n, tile=tile); two.playout(37, seed=seed, no_eye=no_eye); two.playout(plies - 37, seed=seed, no_eye=no_eye)
    ok = ok and torch.equal(two.states, env.states)
    print("parity n=%d G=%d tile=%d no_eye=%d: %s (%d plies)" % (n, G, tile, no_eye, "OK" if ok else "FAIL", nn))
    return ok

def timing(n, G, tile, reps=5):
    env = BatchedGoEnv(G, board_size=n, tile=tile)
    cnt = torch.zeros([1], dtype=torch.int64, device="cuda")
    best = 1e30
    for r in range(reps + 2):
        env.reset(); cnt.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); env.playout(env.max_step + 1, seed=1 + r, counter=cnt); b.record()
        torch.cuda.synchronize()
        if r >= 2:
            best = min(best, a.elapsed_time(b))
    return best, int(cnt.item()) / best * 1e3

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        allok = True
        for n, G, tile, ne in ((9, 4096, 8 | 128, False), (9, 4099, 16 | 128, True), (9, 1000, 32 | 128, False), (9, 4096, 1 | 128, False),
                               (9, 777, 1 | 128, True), (9, 4096, 8 | 128, True), (9, 4096, 4 | 128, False), (9, 4101, 4 | 128, True), (19, 130, 4 | 128, False), (19, 256, 8 | 128, False), (19, 100, 32 | 128, True),
                               (19, 256, 1 | 128, False), (9, 4096, 0, False)):
            allok &= parity(n, G, tile, ne)
        print("ALL PARITY", "OK" if allok else "FAILED")
    if what in ("all", "time"):
        for n, Gs, tiles in ((9, (2048, 4096, 8192, 12288, 16384, 32768, 65536), (4 | 128, 8 | 128, 16 | 128, 1 | 128)),
                             (19, (512, 1024, 2048, 4096, 8192, 16384), (4 | 128, 8 | 128, 16 | 128, 32 | 128, 1 | 128))):
            for G in Gs:
                row = []
                for tile in tiles:
                    if G >= 65536 and (tile & 63) >= 16 and tile > 0:
                        row.append("%12s" % "-"); continue
                    ms, rate = timing(n, G, tile, 3)
                    row.append("%5.3fms %.2e" % (ms, rate))
                print("n=%d G=%7d | " % (n, G) + " | ".join("t%d:%s" % (t, r) for t, r in zip(tiles, row)), flush=True)
