"""GPU tests of the output side of the path: trajectory assembly kernel and device Dirichlet noise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _reference_rows(obs, pi, player, length, winner, terr, fin, n):
    """self_play.py:931-965 restated with numpy, game by game / position by position / rot90 then fliplr."""
    rows = []
    for g in range(obs.shape[0]):
        if fin is not None and not fin[g]:
            continue
        for t in range(int(length[g])):
            o, p, pl = obs[g, t], pi[g, t], int(player[g, t])
            z = 1.0 if pl == int(winner[g]) else -1.0
            own = terr[g] if pl == 1 else -terr[g]
            for i in (1, 2, 3, 4):
                ep = np.rot90(p[:-1].reshape(n, n), i)
                eo = np.array([np.rot90(s, i) for s in o]); ew = np.rot90(own.reshape(n, n), i)
                rows.append((eo, np.append(ep.flatten(), p[-1]), z, ew.flatten()))
                rows.append((np.array([np.fliplr(s) for s in eo]), np.append(np.fliplr(ep).flatten(), p[-1]), z, np.fliplr(ew).flatten()))
    return rows


@pytest.mark.parametrize("n,G,T", [(9, 7, 13), (19, 3, 5)])
def test_assembly_kernel_matches_the_reference_loop(n, G, T):
    from transgo_b200.selfplay import assemble_trajectories, augment8, augment8_rows
    rng = np.random.default_rng(n)
    P, A = n * n, n * n + 1
    obs = rng.random((G, T, 10, n, n)).astype(np.float32); pi = rng.random((G, T, A)).astype(np.float32)
    player = rng.integers(1, 3, (G, T)).astype(np.int32); length = rng.integers(0, T + 1, G).astype(np.int64)
    winner = rng.integers(1, 3, G).astype(np.uint8); terr = rng.integers(-1, 2, (G, P)).astype(np.float32)
    fin = rng.random(G) < 0.7
    fin[0] = True
    d = lambda a: torch.from_numpy(a).cuda()
    for mask in (None, fin):
        exp = _reference_rows(obs, pi, player, length, winner, terr, mask, n)
        o, p, z, w, rows = assemble_trajectories(d(obs), d(pi), d(player), d(length), d(winner), d(terr), fin=None if mask is None else d(mask), augment=True)
        assert rows == len(exp) == o.shape[0]
        assert np.array_equal(o.cpu().numpy(), np.array([r[0] for r in exp])) and np.array_equal(p.cpu().numpy(), np.array([r[1] for r in exp], np.float32))
        assert np.array_equal(z.cpu().numpy(), np.array([r[2] for r in exp], np.float32)) and np.array_equal(w.cpu().numpy(), np.array([r[3] for r in exp], np.float32))
        # as played (no symmetries): every eighth row of the reference expansion is the 4-quarter-turn identity
        o1, p1, z1, w1, rows1 = assemble_trajectories(d(obs), d(pi), d(player), d(length), d(winner), d(terr), fin=None if mask is None else d(mask), augment=False)
        assert rows1 * 8 == rows and torch.equal(o1, o[6::8]) and torch.equal(p1, p[6::8]) and torch.equal(w1, w[6::8])
        # labelled rows expanded at the receiver: same result as the torch op chain
        a = augment8(o1, p1, z1, w1, n); b = augment8_rows(o1, p1, z1, w1, n)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        assert torch.equal(b[0], o) and torch.equal(b[3], w)


def test_assembly_writes_straight_into_a_ring_with_wrap():
    from transgo_b200.replay import ReplayRing
    from transgo_b200.selfplay import assemble_trajectories
    n, G, T = 9, 5, 6
    rng = np.random.default_rng(1)
    d = lambda a: torch.from_numpy(a).cuda()
    obs = d(rng.random((G, T, 10, n, n)).astype(np.float32)); pi = d(rng.random((G, T, 82)).astype(np.float32))
    player = d(rng.integers(1, 3, (G, T)).astype(np.int32)); length = d(np.array([6, 2, 0, 5, 1], np.int64))
    winner = d(rng.integers(1, 3, G).astype(np.uint8)); terr = d(rng.integers(-1, 2, (G, 81)).astype(np.float32))
    ring = ReplayRing(200, device="cuda")
    ring.obs.fill_(7.0)
    ring.index = 150                                                  # 14 positions x 8 = 112 rows: wraps to slot 62
    fresh = assemble_trajectories(obs, pi, player, length, winner, terr, augment=True)
    rows = ring.append_games(obs, pi, player, length, winner, terr, augment=True)
    assert rows == 112 == fresh[4] and ring.full and ring.index == (150 + 112) % 200
    pos = (torch.arange(112, device="cuda") + 150) % 200
    assert torch.equal(ring.obs[pos], fresh[0]) and torch.equal(ring.pi[pos], fresh[1])
    assert torch.equal(ring.z[pos], fresh[2]) and torch.equal(ring.own[pos], fresh[3])
    assert bool((ring.obs[62:150] == 7.0).all())                      # rows outside the written window are untouched


def test_device_dirichlet_noise():
    """tg_dirichlet: rows sum to one over the first K entries, zeros beyond, moments of a symmetric Dirichlet(0.03), and the same
    draws for a shard of the games (counter-based: keyed by the global game id)."""
    from transgo_b200 import _lib
    L = _lib.lib()
    G, A, alpha = 4096, 82, 0.03
    k = torch.full([G], 60, dtype=torch.int32, device="cuda"); k[::2] = 82
    noise = torch.empty([G, A], dtype=torch.float64, device="cuda")
    _lib.check(L.tg_dirichlet(_lib.ptr(noise), _lib.ptr(k), G, A, alpha, 1234, 1, 0, _lib.stream_ptr()))
    x = noise.cpu().numpy()
    assert np.allclose(x.sum(1), 1.0, atol=1e-12) and (x >= 0).all() and not x[1::2, 60:].any()
    # reference values from numpy.random.default_rng(0).dirichlet([0.03] * K, 40000): P(max > 0.5), mean #{x_i > 0.01}
    expect = {82: (0.307, 7.94), 60: (0.457, 6.45)}
    for K, rows in ((82, x[::2, :82]), (60, x[1::2, :60])):
        assert abs(rows.mean() - 1.0 / K) < 1e-9                      # exact by normalisation
        m2 = (alpha + 1) / (K * (K * alpha + 1))                      # E[x_i^2] of a symmetric Dirichlet
        assert abs((rows ** 2).mean() / m2 - 1) < 0.05
        assert abs((rows.max(1) > 0.5).mean() - expect[K][0]) < 0.04   # the shape of the distribution, not just two moments
        assert abs((rows > 0.01).sum(1).mean() - expect[K][1]) < 0.3
    shard = torch.empty([8, A], dtype=torch.float64, device="cuda")
    _lib.check(L.tg_dirichlet(_lib.ptr(shard), _lib.ptr(k[100:108].contiguous()), 8, A, alpha, 1234, 1, 100, _lib.stream_ptr()))
    assert torch.equal(shard, noise[100:108])
    other = torch.empty([8, A], dtype=torch.float64, device="cuda")
    _lib.check(L.tg_dirichlet(_lib.ptr(other), _lib.ptr(k[100:108].contiguous()), 8, A, alpha, 1234, 2, 100, _lib.stream_ptr()))
    assert not torch.equal(other, shard)                              # a new move draws new noise
