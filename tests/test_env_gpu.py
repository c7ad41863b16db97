"""GPU parity tests of the board kernels, all through the C ABI (include/transgo_b200.h).

Checker = oracle/ (C restatement, pinned to the reference by tests/test_oracle.py) and the reference's own golden
traces in tests/golden/.  Bar: bit-exact (integer / 0-1 float work).
"""
import numpy as np
import pytest

from helpers import load_golden, replay_golden, playout_action

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _oracle(n, **kw):
    from oracle.go_oracle import OracleEnv
    return OracleEnv(n, **kw)


def _unpack_states(env):
    return env.states.cpu().numpy()


# ------------------------------------------------------------------ the reference-named single-state symbols
@pytest.mark.parametrize("n,games", [(9, 12), (19, 1)])
def test_compat_symbols_reproduce_reference_golden(oracle_built, n, games, capfd):
    """Unmodified-environment.py view: GoEnv -> Reset/Step/Encode/... -> CUDA, against traces the reference produced."""
    from transgo_b200.environment import GoEnv
    t = load_golden(n)
    env = GoEnv(board_size=n)
    assert env.max_step == int(t["max_step"])
    checked = replay_golden(env, t, games)
    assert checked > 100 * games
    capfd.readouterr()


def test_compat_misc_semantics(oracle_built, capfd):
    from transgo_b200.environment import GoEnv
    env = GoEnv(board_size=9)
    o = _oracle(9)
    s, done = env.reset()
    assert done is False and env.justStarted(s) and env.getPlayer(s) == 1 and env.getStep(s) == 1
    assert env.getScore(s) == -7.5                                   # stone-less board (board.cc:932-935)
    sc, terr = env.getScoreAndTerritory(s)
    assert sc == -7.5 and not terr.any()
    # occupied point is rejected, state unchanged, done False (go_env.cc:75-79)
    s1, d = env.step(s, 40)
    s2, d2 = env.step(s1, 40)
    assert d2 is False and bytes(s2.raw[:48]) == bytes(s1.raw[:48])
    assert env.checkAction(s1, 40) is False and env.checkAction(s1, 41) is True and env.checkAction(s1, -1) is True
    # pass / pass ends the game; stepping a finished game returns True and keeps the state (go_env.cc:52-55)
    p1, d = env.step(s1, 81); assert d is False
    p2, d = env.step(p1, 81); assert d is True and env.isTerminated(p2)
    p3, d = env.step(p2, 3); assert d is True and bytes(p3.raw[:48]) == bytes(p2.raw[:48])
    # resign (-2) ends at once (board.cc:656-661)
    r1, d = env.step(s1, -2); assert d is True
    # sub-board crops (board.cc:1166-1271)
    enc = np.arange(810, dtype=np.float32).reshape(10, 9, 9)
    assert np.array_equal(env.subEncode(enc), o.subEncode(enc, 7, 4))
    assert env.action_to_location(40) == [4, 4] and env.location_to_action([4, 4]) == 40
    env.show(s1)
    out = capfd.readouterr().out
    assert "Fail to Step: invalid action" in out and "Fail to Step: game is done!!" in out and "next_player" in out


# ------------------------------------------------------------------ batched kernels
def _replay_trace_with_oracle(n, trace, game_ids, seed, max_step, no_eye=False):
    """Replays device-chosen actions through the oracle, checking the move rule itself on the way."""
    o = _oracle(n, max_step=max_step)
    P = n * n
    finals = []
    for col, g in enumerate(game_ids):
        s, _ = o.reset()
        for tply in range(trace.shape[0]):
            a = int(trace[tply, col])
            if o.isTerminated(s):
                assert a == -3
                continue
            pts = [x for x in (o.getLegalNoEye(s) if no_eye else o.legal_c(s)) if x != P]
            assert a == playout_action(seed, g, o.getStep(s), pts, P), (g, tply, a)
            s, _ = o.step(s, a)
        finals.append(s)
    return o, finals


@pytest.mark.parametrize("n,G,tile,no_eye", [(9, 4096, 32, False), (9, 4096, 8, False), (9, 4096, 1, False),
                                             (9, 1024, 32, True), (19, 256, 32, False), (19, 256, 1, False),
                                             (9, 4096, 16, False), (9, 4099, 16 | 64, False), (9, 1001, 8 | 64, True), (19, 67, 4 | 64, False), (19, 200, 8 | 64, False), (9, 4096, -1, False), (9, 1000, -1, True), (9, 1, -1, False),
                                             (19, 256, -1, False), (19, 37, -1, True),
                                             (9, 4096, 8 | 128, False), (9, 999, 16 | 128, True), (9, 3000, 1 | 128, True), (19, 300, 32 | 128, False),
                                             (19, 70, 8 | 128, True), (19, 33, 1 | 128, False), (9, 4096, 4 | 128, False), (9, 1003, 4 | 128, True), (19, 150, 4 | 128, True)])
def test_playout_matches_oracle(oracle_built, n, G, tile, no_eye):
    """G concurrent random playouts in one launch; a sample of games is replayed on the CPU."""
    from transgo_b200.environment import BatchedGoEnv
    seed = 0x7261
    env = BatchedGoEnv(G, board_size=n, tile=tile)
    plies = env.max_step + 2
    trace = torch.full([plies, G], -9, dtype=torch.int16, device="cuda")
    counter = torch.zeros([1], dtype=torch.int64, device="cuda")
    env.playout(plies, seed=seed, no_eye=no_eye, trace=trace, counter=counter)
    torch.cuda.synchronize()
    tr = trace.cpu().numpy()
    assert int(counter.item()) == int((tr != -3).sum())
    meta = {k: v.cpu().numpy() for k, v in env.meta().items()}
    assert meta["terminated"].all()
    sample = np.unique(np.concatenate([np.arange(0, min(G, 24)), np.random.default_rng(1).integers(0, G, 40 if n == 9 else 6)]))
    o, finals = _replay_trace_with_oracle(n, tr[:, sample], sample, seed, env.max_step, no_eye)
    sc, terr = env.score(want_territory=True)
    planes = env.encode()
    legal = env.unpack_legal(env.legal_mask()).cpu().numpy()
    noeye = env.unpack_legal(env.legal_mask(no_eye=True)).cpu().numpy()
    torch.cuda.synchronize()
    sc, terr, planes = sc.cpu().numpy(), terr.cpu().numpy(), planes.cpu().numpy()
    P = n * n
    for col, g in enumerate(sample):
        s = finals[col]
        osc, oterr = o.getScoreAndTerritory(s)
        assert sc[g] == osc and np.array_equal(terr[g], oterr), g
        assert np.array_equal(planes[g], o.encode(s)), g
        acts = np.asarray(o.getLegalAction(s)); row = np.zeros(P + 1, bool); row[acts] = True
        assert np.array_equal(legal[g], row), g
        ne = [x for x in o.getLegalNoEye(s) if x != P] or [P]
        row2 = np.zeros(P + 1, bool); row2[ne] = True
        assert np.array_equal(noeye[g], row2), g
        assert meta["step"][g] == o.getStep(s) and meta["player"][g] == o.getPlayer(s)


@pytest.mark.parametrize("n,G,tile", [(9, 96, 32), (9, 96, 8), (9, 96, 1), (19, 8, 32), (9, 96, "inc"), (9, 131, "inc"), (19, 11, "inc")])
def test_step_kernel_per_ply_against_oracle(oracle_built, n, G, tile):
    """Per ply: tg_env_step (tile = lanes per game) / tg_env_step_inc (tile = "inc": the incremental engine with the atari set kept
    beside the positions) with oracle-chosen (sometimes illegal / pass / resign) actions; legal mask, done, accepted, planes, score
    and territory must match the oracle after every move."""
    from transgo_b200.environment import BatchedGoEnv
    rng = np.random.default_rng(11)
    env = BatchedGoEnv(G, board_size=n, step_engine="incremental") if tile == "inc" else BatchedGoEnv(G, board_size=n, tile=tile, step_engine="analysis")
    o = _oracle(n)
    P = n * n
    states = [o.reset()[0] for _ in range(G)]
    nply = 140 if n == 9 else 60
    for ply in range(nply):
        acts = np.zeros(G, np.int16)
        for g in range(G):
            s = states[g]
            la = list(o.getLegalAction(s))
            r = rng.random()
            if r < 0.04:
                a = int(rng.integers(0, P))            # may be illegal
            elif r < 0.10:
                a = P if rng.random() < 0.5 else -1     # both spellings of PASS
            elif r < 0.102:
                a = -2                                  # resign
            else:
                a = int(la[rng.integers(0, len(la))])
            acts[g] = a
        legal, done = env.step(torch.from_numpy(acts).cuda())
        lg = env.unpack_legal(legal).cpu().numpy(); dn = done.cpu().numpy(); acc = env.accepted.cpu().numpy()
        check_planes = ply % 7 == 0
        if check_planes:
            planes = env.encode().cpu().numpy()
            sc, terr = env.score(want_territory=True); sc = sc.cpu().numpy(); terr = terr.cpu().numpy()
        for g in range(G):
            before = states[g]
            was_term = o.isTerminated(before)
            s2, d = o.step(before, int(acts[g]))
            assert bool(dn[g]) == d, (ply, g, acts[g])
            changed = o.getStep(s2) != o.getStep(before)
            assert bool(acc[g]) == changed, (ply, g, acts[g], was_term)
            states[g] = s2
            row = np.zeros(P + 1, bool); row[np.asarray(o.getLegalAction(s2))] = True
            assert np.array_equal(lg[g], row), (ply, g)
            if check_planes:
                assert np.array_equal(planes[g], o.encode(s2)), (ply, g)
                osc, oterr = o.getScoreAndTerritory(s2)
                assert sc[g] == osc and np.array_equal(terr[g], oterr), (ply, g)
        meta = {k: v.cpu().numpy() for k, v in env.meta().items()}
        for g in range(G):
            s = states[g]
            assert (meta["step"][g], meta["player"][g], meta["terminated"][g]) == (s.step_count, s.next_player, s.terminated)
            assert (meta["last1"][g], meta["last2"][g]) == (s.last1, s.last2)
            if s.ko_age == 0:
                assert (meta["ko_loc"][g], meta["ko_color"][g], meta["ko_age"][g]) == (s.ko_loc, s.ko_color, 0)


def test_tiles_agree_bit_for_bit_at_full_size():
    """Size-independent property at the BASELINE C2 size and beyond: the tile widths (32, 8, 1 lanes per game), the lock-step variant
    (8 | 64), the row-per-lane layout (-1, a different formulation of the rules) and the incremental engine (n | 128, a third
    formulation: dense boards + persistent atari set) are different programs and must produce identical final states, scores and planes."""
    from transgo_b200.environment import BatchedGoEnv
    G = 65536
    ref = None
    for tile in (32, 8, 1, -1, 8 | 64, 8 | 128, 4 | 128, 1 | 128, 0):
        env = BatchedGoEnv(G, board_size=9, tile=tile)
        env.playout(122, seed=99)
        out = (env.states.clone(), env.score(), env.encode(), env.legal_mask().clone())
        torch.cuda.synchronize()
        if ref is None:
            ref = out
        else:
            for a, b in zip(ref, out):
                assert torch.equal(a, b)
    st, sc, planes, _ = ref
    # every game ran to termination; area scores are odd multiples of .5 within the board
    assert bool(((st[-1][:, 3] & 1) == 1).all())
    assert bool((sc.abs() <= 81 + 7.5).all()) and bool(((sc * 2) % 2 == 1).all())
    # planes are exactly 0/1 and the six stone planes partition the stones
    assert bool(((planes == 0) | (planes == 1)).all())
    assert bool((planes[:, :6].sum(1) <= 1).all())


def test_playout_is_resumable_and_deterministic():
    """Playing 50 + 72 plies in two launches equals 122 plies in one (the hash is keyed by step_count, not launch)."""
    from transgo_b200.environment import BatchedGoEnv
    a = BatchedGoEnv(2048, tile=32); a.playout(122, seed=5)
    b = BatchedGoEnv(2048, tile=32); b.playout(50, seed=5); b.playout(72, seed=5)
    c = BatchedGoEnv(1024, tile=32); c.playout(122, seed=5, game_id0=1024)
    d = BatchedGoEnv(2048, tile=-1); d.playout(31, seed=5); d.playout(91, seed=5)
    e = BatchedGoEnv(2048, tile=8 | 128); e.playout(45, seed=5); e.playout(1, seed=5); e.playout(76, seed=5)   # atari set rebuilt per launch
    f = BatchedGoEnv(2048, tile=1 | 128); f.playout(77, seed=5); f.playout(45, seed=5)
    torch.cuda.synchronize()
    assert torch.equal(a.states, b.states) and torch.equal(a.states, d.states)
    assert torch.equal(a.states, e.states) and torch.equal(a.states, f.states)
    assert torch.equal(a.states[:, 1024:], c.states)      # sharding by game id gives the same games (multi-GPU rule)


def test_encode_writes_in_place_into_a_larger_tensor():
    from transgo_b200.environment import BatchedGoEnv
    env = BatchedGoEnv(37, tile=32)                       # odd count exercises the unaligned tail
    env.playout(60, seed=3)
    big = torch.full([2, 37, 10, 9, 9], 7.0, device="cuda")
    env.encode(out=big[1])
    ref = env.encode()
    torch.cuda.synchronize()
    assert torch.equal(big[1], ref) and bool((big[0] == 7.0).all())


# ------------------------------------------------------------------ the headline launch, every game
def _headline_launch(G, tile, no_eye, seed, n=9):
    from transgo_b200.environment import BatchedGoEnv
    env = BatchedGoEnv(G, board_size=n, tile=tile)
    plies = env.max_step + 2
    trace = torch.full([plies, G], -9, dtype=torch.int16, device="cuda")
    counter = torch.zeros([1], dtype=torch.int64, device="cuda")
    env.playout(plies, seed=seed, no_eye=no_eye, trace=trace, counter=counter)
    torch.cuda.synchronize()
    return env, trace.cpu().numpy(), int(counter.item())


@pytest.mark.parametrize("G,tile,no_eye", [(4096, 0, False), (4096, 0, True), (32768, 0, False), (4096, 8 | 128, False), (4096, 4 | 64, False),
                                           (4099, 16 | 64, True)])
def test_headline_launch_replayed_on_every_game(oracle_built, G, tile, no_eye):
    """BASELINE configs[1] exactly as bench.py launches it (BatchedGoEnv(4096, tile=0): the library's own layout choice),
    with EVERY game replayed ply by ply on the C oracle (move rule, then final planes / score / territory / legal masks /
    meta of all games), not a sample."""
    seed = 0x7261
    env, tr, played = _headline_launch(G, tile, no_eye, seed)
    o = _oracle(9, max_step=env.max_step)
    n, finals = o.replay_playout(tr, seed, 0, no_eye)
    assert n == played == int((tr != -3).sum())
    planes, score, terr, legal, noeye = o.batch_observe(finals)
    sc, te = env.score(want_territory=True)
    assert np.array_equal(sc.cpu().numpy(), score) and np.array_equal(te.cpu().numpy(), terr)
    assert np.array_equal(env.encode().cpu().numpy(), planes)
    assert np.array_equal(env.unpack_legal(env.legal_mask()).cpu().numpy(), legal)
    assert np.array_equal(env.unpack_legal(env.legal_mask(no_eye=True)).cpu().numpy(), noeye)
    meta = {k: v.cpu().numpy() for k, v in env.meta().items()}
    assert meta["terminated"].all()
    assert np.array_equal(meta["step"], np.array([f.step_count for f in finals]))
    assert np.array_equal(meta["player"], np.array([f.next_player for f in finals]))
    assert np.array_equal(meta["last1"], np.array([f.last1 for f in finals]))
    assert np.array_equal(meta["last2"], np.array([f.last2 for f in finals]))


def test_headline_launch_against_the_reference_engine_itself():
    """The same launch checked against the UNMODIFIED reference library (oracle/_ref/go_env9.so, compiled from
    /root/reference by oracle/Makefile; it travels to the GPU box): every ply of every game goes through the reference's
    own getLegalAction / Step, the final positions through its Encode / getTerritory."""
    import ctypes
    from helpers import mix32
    from oracle.go_oracle import RefEnv, have_ref
    if not have_ref(9):
        pytest.skip("oracle/_ref/go_env9.so not built (needs /root/reference at build time)")
    seed, G = 0x7261, 4096
    env, tr, played = _headline_launch(G, 0, False, seed)
    ref = RefEnv(9, max_step=env.max_step)
    L = ref.lib
    plies = tr.shape[0]
    rnd = mix32(seed, np.arange(G, dtype=np.uint64)[:, None], np.arange(plies + 2, dtype=np.uint64)[None, :])   # [G, step]
    buf = np.zeros(82, dtype=np.int32)
    bufp = buf.ctypes.data
    planes_g = env.encode().cpu().numpy()
    sc_g, te_g = env.score(want_territory=True)
    sc_g, te_g = sc_g.cpu().numpy(), te_g.cpu().numpy()
    enc = np.zeros([10, 9, 9], dtype=np.float32); terr = np.zeros([81], dtype=np.float32)
    a_blob, b_blob = ref._new(), ref._new()
    total = 0
    for g in range(G):
        L.Reset(a_blob)
        done = False
        col = tr[:, g]
        for t in range(plies):
            if done:
                assert col[t] == -3, (g, t)
                continue
            k = L.getLegalAction(a_blob, bufp) - 1                 # the C list always ends with PASS (go_env.cc:170)
            step = t + 1                                           # step_count is 1-based and every ply is accepted
            want = 81 if k == 0 else int(buf[(int(rnd[g, step]) * k) >> 32])
            assert col[t] == want, (g, t, want, int(col[t]))
            done = bool(L.Step(a_blob, b_blob, want))
            a_blob, b_blob = b_blob, a_blob
            total += 1
        if g % 8 == 0:                                             # observables of the final position, every 8th game
            L.Encode(a_blob, enc.ctypes.data)
            sc = L.getTerritory(a_blob, terr.ctypes.data)
            assert np.array_equal(planes_g[g], enc) and sc_g[g] == sc and np.array_equal(te_g[g], terr), g
    assert total == played


@pytest.mark.parametrize("n,G,tile", [(9, 4096, 0), (9, 20000, 0), (19, 300, 0), (9, 513, 8 | 64), (19, 64, -1)])
def test_fused_reset_playout_score_equals_three_launches(n, G, tile):
    """tg_env_playout_scored (reset + playout + final score in ONE launch of the incremental kernel; separate launches behind the
    same entry point for the other layouts) against reset(), playout(), score()."""
    from transgo_b200.environment import BatchedGoEnv
    a = BatchedGoEnv(G, board_size=n, tile=tile)
    a.playout(30, seed=3)                                             # dirty states: from_reset must ignore them
    sa = torch.empty([G], dtype=torch.float32, device="cuda")
    ta = torch.full([a.max_step + 1, G], -9, dtype=torch.int16, device="cuda")
    a.playout(a.max_step + 1, seed=11, trace=ta, from_reset=True, score_out=sa)
    b = BatchedGoEnv(G, board_size=n, tile=tile)
    tb = torch.full([b.max_step + 1, G], -9, dtype=torch.int16, device="cuda")
    b.reset(); b.playout(b.max_step + 1, seed=11, trace=tb)
    torch.cuda.synchronize()
    assert torch.equal(a.states, b.states) and torch.equal(ta, tb) and torch.equal(sa, b.score())
    # scoring alone, mid-game (not from reset): 40 more plies of a running game + its score
    c = BatchedGoEnv(G, board_size=n, tile=tile); c.playout(25, seed=5)
    d = BatchedGoEnv(G, board_size=n, tile=tile); d.playout(25, seed=5)
    sc = torch.empty([G], dtype=torch.float32, device="cuda")
    c.playout(40, seed=5, score_out=sc); d.playout(40, seed=5)
    assert torch.equal(c.states, d.states) and torch.equal(sc, d.score())
    e = BatchedGoEnv(G, board_size=n, tile=tile)
    se = torch.empty([G], dtype=torch.float32, device="cuda")
    e.playout(0, from_reset=True, score_out=se)                        # stone-less boards score -komi (board.cc:932-935)
    assert bool((se == -7.5).all())


def test_playout_on_pinned_host_buffers_equals_the_device_run():
    """zero-copy route: states, move trace and scores in pinned host memory, read / written by the kernel across PCIe; every
    byte must equal the device-resident run (9x9 headline batch and a 19x19 batch)."""
    import torch
    from transgo_b200.environment import BatchedGoEnv
    for n, G in ((9, 4096), (19, 96)):
        env = BatchedGoEnv(G, board_size=n)
        T = env.max_step + 1
        d_trace = torch.full([T, G], -9, dtype=torch.int16, device="cuda")
        d_score = torch.empty([G], dtype=torch.float32, device="cuda")
        env.playout(T, seed=77, trace=d_trace, score_out=d_score)
        h_states = torch.zeros([env.SQ, G, 4], dtype=torch.int32).pin_memory()
        h_states.copy_(BatchedGoEnv(G, board_size=n).states)
        h_trace = torch.full([T, G], -9, dtype=torch.int16).pin_memory()
        h_score = torch.zeros([G], dtype=torch.float32).pin_memory()
        env.playout(T, seed=77, trace=h_trace, score_out=h_score, states=h_states)
        torch.cuda.synchronize()
        assert torch.equal(h_states, env.states.cpu()) and torch.equal(h_trace, d_trace.cpu()) and torch.equal(h_score, d_score.cpu())
        # separate input / output buffers (tg_env_playout_io): the input stays untouched, the output equals the in-place result
        h_in = torch.zeros([env.SQ, G, 4], dtype=torch.int32).pin_memory()
        h_in.copy_(BatchedGoEnv(G, board_size=n).states)
        keep = h_in.clone()
        h_out = torch.zeros_like(h_in).pin_memory()
        h_trace.fill_(-9); h_score.zero_()
        env.playout(T, seed=77, trace=h_trace, score_out=h_score, states=h_out, states_in=h_in)
        torch.cuda.synchronize()
        assert torch.equal(h_in, keep) and torch.equal(h_out, h_states) and torch.equal(h_trace, d_trace.cpu()) and torch.equal(h_score, d_score.cpu())
        for tile in (-1, 8 | 64):                                                        # layouts that play in place: copy, then play
            other = BatchedGoEnv(G, board_size=n, tile=tile)
            d_out = torch.zeros_like(env.states)
            other.playout(T, seed=77, states=d_out, states_in=h_in)
            torch.cuda.synchronize()
            assert torch.equal(d_out.cpu(), h_states), tile
        with pytest.raises(AssertionError):
            env.playout(T, states=torch.zeros([env.SQ, G, 4], dtype=torch.int32))       # pageable host memory is refused


def test_step_graph_replays_equal_eager_steps():
    """BatchedGoEnv.make_step_graph: step (+ encode) captured once and replayed per ply must give what the eager calls give."""
    import torch
    from transgo_b200.environment import BatchedGoEnv
    G = 512
    a = BatchedGoEnv(G, board_size=9); b = BatchedGoEnv(G, board_size=9)
    planes = torch.zeros([G, 10, 9, 9], dtype=torch.float32, device="cuda")
    step = b.make_step_graph(planes=planes)
    assert torch.equal(a.states, b.states)                       # building the graph leaves the position alone
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    for ply in range(40):
        legal = a.unpack_legal(a.legal_mask())                   # [G, 82]
        r = torch.rand([G, 82], device="cuda", generator=gen) * legal
        acts = r.argmax(1).to(torch.int16)
        if ply % 7 == 3:
            acts[::5] = 0                                        # some illegal / occupied points: refused, state unchanged
        la, da = a.step(acts)
        lb, db = step(acts)
        assert torch.equal(a.states, b.states) and torch.equal(la, lb) and torch.equal(da, db), ply
        assert torch.equal(planes, a.encode()), ply


def test_incremental_step_survives_positions_changed_behind_its_back():
    """The atari records of tg_env_step_inc are a cache keyed by a digest of the stones: after the positions were replaced by
    something else (a playout, a copy from another env) the next step must rebuild them and agree with the analysis kernel."""
    import torch
    from transgo_b200.environment import BatchedGoEnv
    for n, G in ((9, 1000), (19, 40)):
        a = BatchedGoEnv(G, board_size=n, step_engine="incremental")
        b = BatchedGoEnv(G, board_size=n, step_engine="analysis")
        gen = torch.Generator(device="cuda"); gen.manual_seed(5)

        def both(acts):
            la, da = a.step(acts); lb, db = b.step(acts)
            assert torch.equal(a.states, b.states) and torch.equal(la, lb) and torch.equal(da, db) and torch.equal(a.accepted, b.accepted)

        def random_legal():
            legal = b.unpack_legal(b.legal_mask())
            return (torch.rand([G, n * n + 1], device="cuda", generator=gen) * legal).argmax(1).to(torch.int16)
        for _ in range(6):
            both(random_legal())
        a.playout(30, seed=9); b.playout(30, seed=9)                      # the records of `a` are stale now
        for _ in range(6):
            both(random_legal())
        other = BatchedGoEnv(G, board_size=n); other.playout(55, seed=4)
        a.states.copy_(other.states); b.states.copy_(other.states)         # stale again, mid-game positions with kos and ataris
        a.atari.random_(0, 2 ** 31 - 1)                                    # and garbage records on top
        for _ in range(10):
            both(random_legal())
