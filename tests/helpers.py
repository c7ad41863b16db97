"""Shared test helpers: golden-trace replay and the numpy mirror of the device hash."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(n):
    z = np.load(os.path.join(GOLDEN, "env_trace_%d.npz" % n))
    return {k: z[k] for k in z.files}


def replay_golden(env, t, max_games=None, check_noeye=True):
    """Replay a reference trace through `env` (any object with the environment.py method names) and require every
    observable to equal what the reference produced.  Returns the number of positions checked."""
    n = int(t["n"]); P = n * n
    legal = np.unpackbits(t["legal"], axis=1)[:, :P + 1]
    noeye = np.unpackbits(t["noeye"], axis=1)[:, :P + 1]
    planes = np.unpackbits(t["planes"], axis=1)[:, :10 * P].reshape(-1, 10, n, n)
    checked = 0
    s = None
    cur_game = -1
    for i in range(len(t["action"])):
        g = int(t["game"][i])
        if max_games is not None and g >= max_games:
            break
        if g != cur_game:
            s, _ = env.reset(); cur_game = g
        acts = env.getLegalAction(s)
        ref_acts = np.nonzero(legal[i][:P])[0]
        if len(ref_acts) == 0:
            ref_acts = np.array([P])
        assert np.array_equal(np.asarray(acts, dtype=np.int64), ref_acts), ("legal", i, g)
        if check_noeye:
            assert np.array_equal(np.asarray(env.getLegalNoEye(s), dtype=np.int64), np.nonzero(noeye[i])[0]), ("noeye", i)
        enc = env.encode(s)
        assert enc.dtype == np.float32 and np.array_equal(enc, planes[i].astype(np.float32)), ("planes", i, g)
        sc, terr = env.getScoreAndTerritory(s)
        assert sc == t["score"][i], ("score", i, sc, t["score"][i])
        assert np.array_equal(terr, t["terr"][i].astype(np.float32)), ("territory", i)
        assert env.getScore(s) == t["score"][i]
        assert env.getPlayer(s) == int(t["player"][i]) and env.getStep(s) == int(t["step"][i]), ("meta", i)
        checked += 1
        a = int(t["action"][i])
        if a == -3:
            continue
        s2, done = env.step(s, a)
        assert bool(done) == bool(t["done"][i]), ("done", i, a)
        assert (env.getStep(s2) != env.getStep(s)) == bool(t["accepted"][i]), ("accepted", i, a)
        s = s2
    return checked


def mix32(seed, a, b):
    """numpy mirror of tg::mix32 (transgo_b200/csrc/tg_board.cuh)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.asarray(a, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + np.asarray(b, dtype=np.uint64) * np.uint64(0xBF58476D1CE4E5B9))
        z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(32)).astype(np.uint64)


def playout_action(seed, game_id, step, legal_points, P):
    """The SURVEY 8d C2 move rule: k-th legal point, k = (mix32 * n_legal) >> 32; PASS iff none."""
    nl = len(legal_points)
    if nl == 0:
        return P
    r = int(mix32(seed, game_id, step))
    return int(legal_points[(r * nl) >> 32])


# ----------------------------------------------------------------------------------------------- MCTS golden
def load_mcts_golden():
    z = np.load(os.path.join(GOLDEN, "mcts_trace_9.npz"))
    return {k: z[k] for k in z.files}


def mcts_golden_games(t):
    """Yields (game id, sims, sharp, row indices) per recorded game."""
    for g in np.unique(t["game"]):
        rows = np.nonzero(t["game"] == g)[0]
        yield int(g), int(t["sims"][rows[0]]), int(t["sharp"][rows[0]]), rows


def golden_hooks(t, rows, game):
    """noise / uniform / tie-break hooks that replay exactly what the reference run consumed."""
    from oracle import provider
    seed = int(t["seed"])
    ties = [0]

    def tie(n):
        i = provider.tie_index(seed, game, ties[0], n)
        ties[0] += 1
        return i
    return dict(noise_fn=lambda m, k: t["noise"][rows[m]][:k], u_fn=lambda m: float(t["u"][rows[m]]), tie_fn=tie), ties


def check_mcts_move_against_golden(t, row, acts, N, W, var, action, pi, done, rtol=1e-5):
    """Visit counts exact; W / var within 1e-5 relative (the north_star tolerance for fp32 Q statistics); the
    reference arm itself is float32, so the oracle is expected to hit them exactly."""
    counts = np.zeros(82, np.int64); counts[acts] = N
    assert np.array_equal(counts, t["counts"][row]), ("visit counts", row)
    w = np.zeros(82, np.float32); w[acts] = W
    v = np.zeros(82, np.float32); v[acts] = var
    np.testing.assert_allclose(w, t["W"][row], rtol=rtol, atol=1e-6)
    np.testing.assert_allclose(v, t["var"][row], rtol=rtol, atol=1e-6)
    assert int(action) == int(t["action"][row]), ("action", row)
    assert np.array_equal(np.asarray(pi, dtype=np.float64), t["pi"][row]), ("pi", row)
    assert bool(done) == bool(t["done"][row])


def torch_hash_provider(seed, n, sharp=3, device="cuda"):
    """Device twin of oracle.provider.hash_eval_batch: exact int64 tensor ops + one IEEE float32 division per output, no
    host round trip -> usable as the evaluator of a CUDA-graphed wave, and bit-identical to the numpy twin."""
    import torch
    from oracle import provider
    A = n * n + 1
    coef, mul, add = (torch.from_numpy(t).to(device) for t in provider.hash_eval_tables(seed, 10 * n * n, A))
    den = torch.full([1], 32832.0, dtype=torch.float32, device=device)   # tensor / python-scalar becomes a multiply by 1/x on CUDA

    def evaluator(obs):
        B = obs.shape[0]
        bits = obs.reshape(B, -1).to(torch.int64)
        h = (bits * coef[None, :]).sum(1) & 0x7FFFFFFF
        r = ((h[:, None] * mul[None, :] + add[None, :]) >> 13) & 0xFFFF
        w = r
        for _ in range(sharp - 1):
            w = (w * r) >> 16
        w = w + 1
        total = w.sum(1).to(torch.float32)
        policy = w.to(torch.float32) / total[:, None]
        v = (((h * 2654435761) >> 9) & 0xFFFF) - 32768
        return policy, v.to(torch.float32) / den
    return evaluator
