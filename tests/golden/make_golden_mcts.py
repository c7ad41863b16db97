"""Generate tests/golden/mcts_trace_9.npz from the UNMODIFIED reference search (self_play.py WP_MCTS driven through
oracle/ref_mcts_harness.py).  Runs only where /root/reference exists; the .npz is committed.

Per game: the search settings and provider seed; per move: the injected Dirichlet noise and uniform draw, and what the
reference produced -- root child visit counts / W / var (float32), root N, pi, the chosen action, done.
Tie-breaks follow oracle.provider.tie_index(seed, game, counter, n) (integer hashing, reproducible anywhere).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import provider  # noqa: E402
from oracle.ref_mcts_harness import ReferenceMCTS  # noqa: E402

GAMES = [  # (game id, num_simulation, provider sharpness, max moves)
    (0, 64, 3, 200), (1, 100, 6, 40), (2, 40, 8, 200), (3, 210, 3, 6),
]
SEED = 77
A = 82


def one_game(game, sims, sharp, max_moves):
    ev = lambda obs: provider.synth_eval_batch(obs, seed=SEED, sharp=sharp)
    ties = [0]

    def tie(n):
        i = provider.tie_index(SEED, game, ties[0], n)
        ties[0] += 1
        return i
    noises, us = [], []

    def noise_fn(m, k):
        x = provider.dirichlet(SEED, game, m, k)
        row = np.zeros(A); row[:k] = x
        noises.append(row)
        return x

    def u_fn(m):
        u = provider.uniform(SEED, game, m)
        us.append(u)
        return u
    ref = ReferenceMCTS(ev, num_simulation=sims, tie_fn=tie, noise_fn=noise_fn, u_fn=u_fn)
    rec = dict(counts=[], W=[], var=[], rootN=[], pi=[], action=[], done=[], nties=[])
    for m in range(max_moves):
        a, pi, _ = ref.get_action_probs()
        acts, N, W, var, _ = ref.root_children()
        c = np.zeros(A, np.int32); c[acts] = N
        w = np.zeros(A, np.float32); w[acts] = W
        v = np.zeros(A, np.float32); v[acts] = var
        done = ref.update_with_action(a)
        rec["counts"].append(c); rec["W"].append(w); rec["var"].append(v); rec["rootN"].append(ref_rootN(ref, acts, N))
        rec["pi"].append(pi); rec["action"].append(a); rec["done"].append(done); rec["nties"].append(ties[0])
        if done:
            break
    n = len(rec["action"])
    return dict(game=np.full(n, game, np.int32), sims=np.full(n, sims, np.int32), sharp=np.full(n, sharp, np.int32),
                move=np.arange(n, dtype=np.int32), noise=np.array(noises), u=np.array(us),
                counts=np.array(rec["counts"]), W=np.array(rec["W"]), var=np.array(rec["var"]),
                rootN=np.array(rec["rootN"], np.int32), pi=np.array(rec["pi"]), action=np.array(rec["action"], np.int16),
                done=np.array(rec["done"], np.bool_), nties=np.array(rec["nties"], np.int32))


def ref_rootN(ref, acts, N):
    return int(N.sum())   # recorded before update_with_action by the caller's ordering; see below


if __name__ == "__main__":
    parts = [one_game(*g) for g in GAMES]
    out = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    out["seed"] = np.int64(SEED)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mcts_trace_9.npz"), **out)
    print("mcts_trace_9: %d moves over %d games, %d tie-breaks" % (len(out["action"]), len(GAMES), int(sum(p["nties"][-1] for p in parts))))
