"""Generate tests/golden/env_trace_{9,19}.npz from the UNMODIFIED reference engine (oracle/_ref, built by
oracle/Makefile from /root/reference/GoEnv/cpp_src).  Runs only where /root/reference exists; the .npz files are
committed so that the oracle and the CUDA kernels can be pinned on machines without the reference.

Per ply of every game the file records what GoEnv/environment.py makes observable:
  action taken, C-level legal list (as an 0/1 row over A actions, PASS column always 1), getLegalNoEye row,
  the ten encode10 planes (bit-packed), getScore, getTerritory (int8), done flag returned by Step, getPlayer, getStep.
Three playout policies are mixed: uniform over environment.py's legal list, uniform over getLegalNoEye, and
"mixed" (random passes + occasional random, possibly illegal, points).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.go_oracle import RefEnv  # noqa: E402


def trace(n, games, seed):
    rng = np.random.default_rng(seed)
    env = RefEnv(n)
    P = n * n
    rec = {k: [] for k in ("game", "action", "legal", "noeye", "planes", "score", "terr", "done", "player", "step", "accepted_state_changed")}
    for g in range(games):
        policy = ("legal", "noeye", "mixed")[g % 3]
        s, _ = env.reset()
        done = False
        while not done:
            lc = env.legal_c(s)
            ne = env.getLegalNoEye(s)
            row = np.zeros(P + 1, np.uint8); row[lc] = 1
            row2 = np.zeros(P + 1, np.uint8); row2[ne] = 1
            sc, terr = env.getScoreAndTerritory(s)
            if policy == "legal":
                acts = env.getLegalAction(s)
            elif policy == "noeye":
                acts = ne
            else:
                acts = list(env.getLegalAction(s)) + ([P] if rng.random() < 0.15 else [])
                if rng.random() < 0.03:
                    acts = [int(rng.integers(0, P))]
            a = int(acts[rng.integers(0, len(acts))])
            rec["game"].append(g); rec["action"].append(a)
            rec["legal"].append(row); rec["noeye"].append(row2)
            rec["planes"].append(np.packbits(env.encode(s).astype(np.uint8).ravel()))
            rec["score"].append(sc); rec["terr"].append(terr.astype(np.int8))
            rec["player"].append(env.getPlayer(s)); rec["step"].append(env.getStep(s))
            s2, done = env.step(s, a)
            rec["done"].append(done)
            rec["accepted_state_changed"].append(env.getStep(s2) != env.getStep(s))
            s = s2
        # final position of the game (action = -3 marks "no move follows")
        lc = env.legal_c(s); ne = env.getLegalNoEye(s)
        row = np.zeros(P + 1, np.uint8); row[lc] = 1
        row2 = np.zeros(P + 1, np.uint8); row2[ne] = 1
        sc, terr = env.getScoreAndTerritory(s)
        rec["game"].append(g); rec["action"].append(-3); rec["legal"].append(row); rec["noeye"].append(row2)
        rec["planes"].append(np.packbits(env.encode(s).astype(np.uint8).ravel()))
        rec["score"].append(sc); rec["terr"].append(terr.astype(np.int8))
        rec["player"].append(env.getPlayer(s)); rec["step"].append(env.getStep(s)); rec["done"].append(True)
        rec["accepted_state_changed"].append(False)
    out = dict(n=np.int32(n), max_step=np.int32(env.max_step), komi=np.float32(env.komi),
               game=np.array(rec["game"], np.int32), action=np.array(rec["action"], np.int16),
               legal=np.packbits(np.array(rec["legal"]), axis=1), noeye=np.packbits(np.array(rec["noeye"]), axis=1),
               planes=np.array(rec["planes"]), score=np.array(rec["score"], np.float32), terr=np.array(rec["terr"]),
               done=np.array(rec["done"], np.bool_), player=np.array(rec["player"], np.uint8),
               step=np.array(rec["step"], np.int32), accepted=np.array(rec["accepted_state_changed"], np.bool_))
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)   # the reference printf()s on illegal moves
    t9 = trace(9, 36, 20261018)
    t19 = trace(19, 3, 20261019)
    os.dup2(saved, 1)
    np.savez_compressed(os.path.join(here, "env_trace_9.npz"), **t9)
    np.savez_compressed(os.path.join(here, "env_trace_19.npz"), **t19)
    for name, t in (("9", t9), ("19", t19)):
        print("env_trace_%s: %d positions, %d rejected moves" % (
            name, len(t["action"]), int((~t["accepted"] & (t["action"] != -3)).sum())))
