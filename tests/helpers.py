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
