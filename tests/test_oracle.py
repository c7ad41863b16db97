"""CPU tests of the checker itself: the C restatement must reproduce the reference's golden traces, and (where the
compiled reference is present) agree with it on fresh random games."""
import os

import numpy as np
import pytest

from helpers import load_golden, replay_golden


@pytest.mark.parametrize("n,games", [(9, None), (19, None)])
def test_oracle_reproduces_reference_golden(oracle_built, n, games):
    from oracle.go_oracle import OracleEnv
    t = load_golden(n)
    env = OracleEnv(n, max_step=int(t["max_step"]), komi=float(t["komi"]))
    checked = replay_golden(env, t, games)
    assert checked == len(t["action"])


def test_golden_has_the_edge_cases():
    t = load_golden(9)
    P = 81
    planes = np.unpackbits(t["planes"], axis=1)[:, :10 * P].reshape(-1, 10, P)
    legal = np.unpackbits(t["legal"], axis=1)[:, :P + 1]
    assert (~t["accepted"] & (t["action"] != -3)).sum() > 0          # rejected (illegal) moves
    assert (t["action"] == P).sum() > 0                               # passes
    assert planes[:, 8].sum() > 0 and planes[:, 9].sum() > 0          # true eyes and living blocks
    assert planes[:, 7].sum() > 0                                     # ko / suicide points
    assert (legal[:, :P].sum(1) == 0).sum() >= 0
    assert (t["score"] == -7.5).sum() > 0                             # stone-less board quirk (board.cc:932-935)
    assert (t["terr"] == 0).any() and (t["terr"] == 1).any() and (t["terr"] == -1).any()


@pytest.mark.parametrize("n,games", [(9, 40), (19, 2)])
def test_oracle_matches_compiled_reference(oracle_built, n, games, capfd):
    from oracle.go_oracle import OracleEnv, RefEnv, have_ref
    if not have_ref(n):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(7 + n)
    o, r = OracleEnv(n), RefEnv(n)
    P = n * n
    for g in range(games):
        so, _ = o.reset(); sr, _ = r.reset()
        done = False
        while not done:
            assert np.array_equal(o.legal_c(so), r.legal_c(sr))
            assert np.array_equal(o.getLegalNoEye(so), r.getLegalNoEye(sr))
            assert np.array_equal(o.encode(so), r.encode(sr))
            a, b = o.getScoreAndTerritory(so), r.getScoreAndTerritory(sr)
            assert a[0] == b[0] and np.array_equal(a[1], b[1])
            assert (o.getPlayer(so), o.getStep(so)) == (r.getPlayer(sr), r.getStep(sr))
            acts = list(o.getLegalAction(so)) + ([P] if rng.random() < 0.1 else [])
            if rng.random() < 0.03:
                acts = [int(rng.integers(0, P))]
            act = int(acts[rng.integers(0, len(acts))])
            so, d1 = o.step(so, act); sr, d2 = r.step(sr, act)
            assert d1 == d2 and o.isTerminated(so) == r.isTerminated(sr)
            done = d1
    capfd.readouterr()


def test_sub_encode_matches_reference(oracle_built):
    from oracle.go_oracle import OracleEnv, RefEnv, have_ref
    o = OracleEnv(9)
    enc = np.arange(10 * 81, dtype=np.float32).reshape(10, 9, 9)
    sub = o.subEncode(enc, 7, 4)
    assert np.array_equal(sub[0], enc[:, :7, :7]) and np.array_equal(sub[1], enc[:, :7, 2:])
    assert np.array_equal(sub[2], enc[:, 2:, :7]) and np.array_equal(sub[3], enc[:, 2:, 2:])
    if have_ref(9):
        assert np.array_equal(RefEnv(9).subEncode(enc, 7, 4), sub)
        assert np.array_equal(RefEnv(9).subEncode(enc, 7, 5), o.subEncode(enc, 7, 5))
