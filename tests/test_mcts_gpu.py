"""GPU parity tests of the MCTS kernels, through the C ABI (tg_mcts_* in include/transgo_b200.h).

Checker = oracle/mcts_oracle.py (numpy restatement pinned bit-exact to the reference WP_MCTS by tests/test_oracle_mcts.py)
and tests/golden/mcts_trace_9.npz produced by the reference itself.  Bar (BASELINE.json north_star): visit counts
bit-exact given identical net outputs, Dirichlet noise and tie-break order; W / var within 1e-5 relative (fp32).
"""
import numpy as np
import pytest

from helpers import load_mcts_golden, mcts_golden_games, check_mcts_move_against_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def cpu_provider(seed, sharp):
    """The deterministic stand-in net (oracle/provider.py) evaluated on the host: identical bits on every arm."""
    from oracle import provider

    def evaluator(obs):
        p, v = provider.synth_eval_batch(obs.cpu().numpy(), seed=seed, sharp=sharp)
        return torch.from_numpy(p).cuda(), torch.from_numpy(v).cuda()
    return evaluator


def make_oracles(G, n, sims, seed, sharp, game_id0=0, max_step=None, tie="hash", readouts=4, eval_fn=None):
    from oracle import provider
    eval_fn = eval_fn or (lambda o: provider.synth_eval_batch(o, seed=seed, sharp=sharp))
    from oracle.go_oracle import OracleEnv
    from oracle.mcts_oracle import OracleMCTS, MCTSConfig
    out = []
    for g in range(G):
        gid = game_id0 + g
        cnt = [0]

        def tie_fn(k, gid=gid, cnt=cnt):
            if tie != "hash":
                return 0
            i = provider.tie_index(seed, gid, cnt[0], k)
            cnt[0] += 1
            return i
        out.append(OracleMCTS(OracleEnv(n, max_step=max_step), eval_fn,
                              MCTSConfig(board_size=n, num_simulation=sims, parallel_readouts=readouts), tie_fn=tie_fn,
                              noise_fn=lambda m, k, gid=gid: provider.dirichlet(seed, gid, m, k),
                              u_fn=lambda m, gid=gid: provider.uniform(seed, gid, m)))
    return out


def compare_batch(eng, oracles, seed, moves, game_id0=0, selfplay=True):
    """Plays `moves` moves on both sides with identical injected randomness and compares every observable."""
    from oracle import provider
    G, A = eng.G, eng.A
    alive = [True] * G
    for move in range(moves):
        noise = np.zeros([G, A])
        for g in range(G):
            if alive[g]:
                k = len(oracles[g].root.acts)
                noise[g, :k] = provider.dirichlet(seed, game_id0 + g, oracles[g].move_idx, k)
        u = torch.tensor([provider.uniform(seed, game_id0 + g, oracles[g].move_idx) for g in range(G)], dtype=torch.float64).cuda()
        actions, pi, obs = eng.get_action_probs(selfplay, noise=torch.from_numpy(noise).cuda() if selfplay else None, u=u)
        counts, W, var, prior = eng.root_stats(full=True)
        counts, W, var = counts.cpu().numpy(), W.cpu().numpy(), var.cpu().numpy()
        root_n = eng.root_n.cpu().numpy()
        pi, obs, actions = pi.cpu().numpy(), obs.cpu().numpy(), actions.cpu().numpy()
        for g in range(G):
            if not alive[g]:
                continue
            o = oracles[g]
            a, opi, oobs = o.get_action_probs(selfplay)
            acts, N, oW, ovar, _ = o.root_children()
            assert np.array_equal(counts[g], o.visit_counts()), ("visit counts", move, g)          # bit-exact
            assert root_n[g] == o.rootN, ("root N", move, g)
            np.testing.assert_allclose(W[g][acts], oW, rtol=1e-5, atol=1e-6)                          # north_star tolerance
            np.testing.assert_allclose(var[g][acts], ovar, rtol=1e-5, atol=1e-6)
            assert np.array_equal(W[g][acts], oW) and np.array_equal(var[g][acts], ovar), "fp32 statistics are expected to be bit-equal"
            assert np.array_equal(pi[g], opi), ("pi", move, g)
            assert np.array_equal(obs[g], oobs), ("root obs", move, g)
            assert int(actions[g]) == a, ("sampled action", move, g)
        done = eng.update_with_action(torch.from_numpy(actions).cuda()).cpu().numpy()
        for g in range(G):
            if alive[g]:
                d = oracles[g].update_with_action(int(actions[g]))
                assert bool(done[g]) == d, ("done", move, g)
                alive[g] = not d
        if not any(alive):
            break
    return alive


@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_reference_golden_trace(oracle_built, which):
    """The reference WP_MCTS's own recorded games: counts / W / var / pi / action / done per move."""
    from transgo_b200.mcts import BatchedMCTS
    t = load_mcts_golden()
    game, sims, sharp, rows = list(mcts_golden_games(t))[which]
    seed = int(t["seed"])
    eng = BatchedMCTS(cpu_provider(seed, sharp), 1, num_simulation=sims, tie_mode="hash", tie_seed=seed, game_id0=game)
    for row in rows:
        noise = torch.from_numpy(t["noise"][row][None, :].copy()).cuda()
        u = torch.tensor([float(t["u"][row])], dtype=torch.float64).cuda()
        actions, pi, _ = eng.get_action_probs(True, noise=noise, u=u)
        counts, W, var, _ = eng.root_stats(full=True)
        acts = np.nonzero(eng.root_stats(full=True)[3][0].cpu().numpy() > -1)[0]    # all actions; zeros where no edge
        done = eng.update_with_action(actions)
        check_mcts_move_against_golden(t, row, acts, counts[0].cpu().numpy()[acts], W[0].cpu().numpy()[acts],
                                       var[0].cpu().numpy()[acts], int(actions[0].item()), pi[0].cpu().numpy(),
                                       bool(done[0].item()))
        if bool(done[0].item()):
            break


@pytest.mark.parametrize("G,sims,sharp,moves", [(8, 48, 4, 12), (4, 210, 3, 3), (6, 32, 9, 30)])
def test_batched_search_matches_oracle(oracle_built, G, sims, sharp, moves):
    from transgo_b200.mcts import BatchedMCTS
    seed = 100 + sims
    eng = BatchedMCTS(cpu_provider(seed, sharp), G, num_simulation=sims, tie_mode="hash", tie_seed=seed)
    compare_batch(eng, make_oracles(G, 9, sims, seed, sharp), seed, moves)


def test_terminal_heavy_games_and_first_tiebreak(oracle_built):
    """Short games (max_step 14) force terminal leaves (scored and backed up at once, self_play.py:638-642), the
    failsafe of 2*readouts attempts, PASS-only nodes and the game-over path; tie_mode 'first' is the other rule."""
    from transgo_b200.mcts import BatchedMCTS
    seed, G, sims = 31, 6, 40
    eng = BatchedMCTS(cpu_provider(seed, 6), G, num_simulation=sims, max_step=14, tie_mode="first")
    alive = compare_batch(eng, make_oracles(G, 9, sims, seed, 6, max_step=14, tie="first"), seed, 20)
    assert not any(alive)                                      # every game reached its end on both sides
    st = eng.check_status()
    assert not (st & 1).any()


def test_game_id_offset_gives_identical_games(oracle_built):
    """Multi-GPU rule: a shard starting at game_id0 plays exactly the games of the full batch."""
    from transgo_b200.mcts import BatchedMCTS
    seed, sims = 7, 24
    full = BatchedMCTS(cpu_provider(seed, 5), 4, num_simulation=sims, tie_seed=seed)
    shard = BatchedMCTS(cpu_provider(seed, 5), 2, num_simulation=sims, tie_seed=seed, game_id0=2)
    compare_batch(shard, make_oracles(2, 9, sims, seed, 5, game_id0=2), seed, 4, game_id0=2)
    compare_batch(full, make_oracles(4, 9, sims, seed, 5), seed, 4)


def test_wp_mcts_facade_has_the_reference_surface(oracle_built):
    """self_play.py:906-926 call sequence on the G = 1 facade, with a model object exposing main_prediction."""
    from oracle import provider
    from oracle.go_oracle import OracleEnv
    from oracle.mcts_oracle import OracleMCTS, MCTSConfig
    from transgo_b200.environment import GoEnv
    from transgo_b200.mcts import WP_MCTS
    seed = 3

    class Model:
        def main_prediction(self, x):
            p, v = provider.synth_eval_batch(x.cpu().numpy(), seed=seed, sharp=5)
            return torch.from_numpy(p).cuda(), torch.from_numpy(v).cuda(), torch.zeros(len(p), 81).cuda()

    class Cfg:
        board_size, num_simulation, c_puct1, c_puct2, wu_loss, parallel_readouts = 9, 30, 3, 0.05, 2, 4
    env = GoEnv(Cfg())
    agent = WP_MCTS(Cfg(), env, Model(), None, tie_mode="first")
    assert str(agent) == "WP_MCTS" and env.getStep(agent.root.state) == 1
    # evaluation mode (no noise, temperature 0.12) is deterministic up to the move draw -> compare visit counts
    ora = OracleMCTS(OracleEnv(9), lambda o: provider.synth_eval_batch(o, seed=seed, sharp=5), MCTSConfig(num_simulation=30),
                     u_fn=lambda m: 0.5)
    a, pi, obs = agent.get_action_probs(is_selfplay=False)
    oa, opi, oobs = ora.get_action_probs(is_selfplay=False, forced_action=a)
    assert np.array_equal(pi, opi) and np.array_equal(obs, oobs) and pi.shape == (82,) and obs.shape == (10, 9, 9)
    assert agent.update_with_action(a) is False and ora.update_with_action(a) is False
    assert env.getStep(agent.root.state) == 2
    # select_action(state) (self_play.py:689-704) searches from an arbitrary position
    s, _ = env.reset()
    for mv in (40, 41, 31):
        s, _ = env.step(s, mv)
    act = agent.select_action(s)
    assert 0 <= act <= 81 and env.checkAction(s, act if act != 81 else -1)


def test_19x19_search_matches_oracle(oracle_built):
    from transgo_b200.mcts import BatchedMCTS
    seed, G, sims = 19, 2, 24
    eng = BatchedMCTS(cpu_provider(seed, 4), G, board_size=19, num_simulation=sims, tie_seed=seed)
    compare_batch(eng, make_oracles(G, 19, sims, seed, 4), seed, 3)


def test_full_size_invariants_with_device_provider():
    """BASELINE C3 size (1024 games x 800 simulations per move) with the device-side synthetic provider: properties
    that hold for any net -- per-game budget met, child visits sum to the root's, arena never overflows, and the run is
    reproducible bit for bit."""
    from transgo_b200 import _lib
    from transgo_b200.mcts import BatchedMCTS
    G, sims = 1024, 800
    L = _lib.lib()

    def run():
        wave = [0]

        def evaluator(obs):
            B = obs.shape[0]
            p = torch.empty([B, 82], dtype=torch.float32, device="cuda")
            v = torch.empty([B], dtype=torch.float32, device="cuda")
            _lib.check(L.tg_synth_provider(9, B, 12345, wave[0], _lib.ptr(p), _lib.ptr(v), _lib.stream_ptr()))
            wave[0] += 1
            return p, v
        eng = BatchedMCTS(evaluator, G, num_simulation=sims, tie_seed=1)
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        noise = eng.make_noise(generator=g)
        u = torch.rand([G], dtype=torch.float64, device="cuda", generator=g)
        actions, pi, _ = eng.get_action_probs(True, noise=noise, u=u)
        counts = eng.counts.clone(); rootn = eng.root_n.clone()
        eng.update_with_action(actions)
        nsims = eng.total_sims()
        return counts, rootn, actions, pi, nsims
    c1, n1, a1, pi1, s1 = run()
    c2, n2, a2, pi2, s2 = run()
    assert torch.equal(c1, c2) and torch.equal(n1, n2) and torch.equal(a1, a2) and s1 == s2
    assert bool((n1 >= sims).all()) and bool((n1 <= sims + 3).all())       # the last wave may overshoot by readouts-1
    assert bool((c1.sum(1) == n1).all())                                    # fresh roots: every backup passes one child
    assert s1 == int(n1.sum().item())
    assert bool((pi1.sum(1) - 1).abs().max() < 1e-12)
    assert bool((c1.gather(1, a1.long()[:, None]) >= 2).all())             # counts of 1 are zeroed before sampling (:668)


def test_selfplay_worker_labels_and_augmentation(oracle_built):
    """self_play.py:917-965 for a batch: positions are labelled with the final winner / ownership from the mover's view and
    expanded x8; the recorded root observations replay on the oracle."""
    from oracle.go_oracle import OracleEnv
    from transgo_b200.selfplay import SelfPlayWorker
    w = SelfPlayWorker(cpu_provider(5, 6), 4, num_simulation=16, max_step=10, tie_seed=5)
    out = w.play_games(augment=False)
    n = out["positions"]
    assert out["obs"].shape == (n, 10, 9, 9) and out["pi"].shape == (n, 82) and out["z"].shape == (n,) and out["own"].shape == (n, 81)
    assert 4 <= n <= 40 and set(out["z"].cpu().tolist()) <= {1.0, -1.0}
    assert bool(((out["pi"].sum(1) - 1).abs() < 1e-5).all())
    winner = out["winner"].cpu().numpy(); score = out["score"].cpu().numpy()
    assert ((score > 0) == (winner == 1)).all()
    # first recorded position of each game is the empty board with black to move: z = +1 iff black won
    first = out["obs"].reshape(n, -1).abs().sum(1) == 0
    zs = out["z"][first].cpu().numpy()
    assert first.sum().item() == 4 and sorted(zs.tolist()) == sorted(np.where(winner == 1, 1.0, -1.0).tolist())
    aug = w.play_games(augment=True)
    assert aug["obs"].shape[0] == 8 * aug["positions"] and aug["pi"].shape == (8 * aug["positions"], 82)


def test_selfplay_round_feeds_the_replay_ring(oracle_built):
    """BASELINE configs[4] in one process: play -> label -> x8 -> ring, counters advance like SharedStorage's
    (shared_storage.py:21-43).  The multi-rank form (NCCL broadcast + gather) is exercised by scripts/selfplay_loop.py
    and, on CPU with gloo, by tests/test_dist_cpu.py."""
    from transgo_b200.net import TransGoNetwork
    from transgo_b200.replay import Pacing, ReplayRing
    from transgo_b200.selfplay import SelfPlayWorker, selfplay_round
    model = TransGoNetwork(board_size=9, num_features=16).cuda().eval()
    w = SelfPlayWorker(cpu_provider(9, 6), 6, num_simulation=12, max_step=8, tie_seed=3)   # 12 < 81 children: some roots end
    ring, pacing = ReplayRing(4096), Pacing()                                               # with every count <= 1 (see mcts.py)
    called = []
    st = selfplay_round(w, model, ring, pacing, refresh=lambda: called.append(1))
    assert called == [1] and st["games"] == 6 and st["appended"] == 8 * st["positions"] == len(ring)
    assert pacing.now_play_steps == st["positions"] and pacing.now_play_games == 6
    obs, pi, z, own = ring.sample(32)
    assert obs.shape == (32, 10, 9, 9) and bool(((pi.sum(1) - 1).abs() < 1e-5).all()) and set(z.cpu().tolist()) <= {1.0, -1.0}
    st2 = selfplay_round(w, model, ring, pacing, augment=False)
    assert len(ring) == st["appended"] + st2["positions"]


def test_graphed_waves_replay_bit_for_bit():
    """BatchedMCTS(graph_waves=True) captures one wave (select, leaves, evaluator, expand) as a CUDA graph; replaying it must
    give exactly the search the eager launches give (same counts, same chosen actions) over several moves."""
    from transgo_b200 import _lib
    from transgo_b200.mcts import BatchedMCTS, _declare
    L = _lib.lib(); _declare(L)
    G = 96

    def make(graph):
        pol = torch.empty([G * 4, 82], dtype=torch.float32, device="cuda"); val = torch.empty([G * 4], dtype=torch.float32, device="cuda")
        pol_r = torch.empty([G, 82], dtype=torch.float32, device="cuda"); val_r = torch.empty([G], dtype=torch.float32, device="cuda")
        wave = torch.zeros([2], dtype=torch.int64, device="cuda")

        def synth(obs):
            p, v = (pol, val) if obs.shape[0] == G * 4 else (pol_r, val_r)
            _lib.check(L.tg_synth_provider_dev(9, obs.shape[0], 5, _lib.ptr(wave), _lib.ptr(p), _lib.ptr(v), _lib.stream_ptr()))
            return p, v
        return BatchedMCTS(synth, G, num_simulation=40, tie_seed=9, graph_waves=graph)
    a, b = make(False), make(True)
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    for move in range(4):
        noise = a.make_noise(generator=g)
        u = torch.rand([G], dtype=torch.float64, device="cuda", generator=g)
        act_a, pi_a, _ = a.get_action_probs(True, noise=noise, u=u)
        act_b, pi_b, _ = b.get_action_probs(True, noise=noise, u=u)
        assert torch.equal(a.counts, b.counts) and torch.equal(act_a, act_b) and torch.equal(pi_a, pi_b), move
        a.update_with_action(act_a); b.update_with_action(act_b)
    assert b._wave_graph is not None and a._wave_graph is None


def test_evaluate_match_alternates_colours_and_counts_wins(oracle_built):
    """self_play.py:987-1040 batched: a provider that always wants to capture the centre vs. the same one -- the bookkeeping
    (colour alternation, winner from the area score with komi, win + lose = games) is what is checked here."""
    from transgo_b200.selfplay import evaluate_match
    out = evaluate_match(cpu_provider(3, 6), cpu_provider(4, 6), n_games=6, num_simulation=24, max_step=12, tie_seed=2)
    assert out["win"] + out["lose"] == 6 and 0.0 <= out["win_ratio"] <= 1.0
    assert out["new_plays_black"].cpu().tolist() == [True, False, True, False, True, False]
    assert set(out["winners"].cpu().tolist()) <= {1, 2} and 1 <= out["plies"] <= 13
    w = out["winners"].cpu().numpy(); nb = out["new_plays_black"].cpu().numpy()
    assert out["win"] == int(((w == 1) & nb).sum() + ((w == 2) & ~nb).sum())


def test_evaluate_match_equals_the_reference_procedure_on_the_oracle(oracle_built):
    """SelfPlay.policy_evaluate (self_play.py:987-1040) move by move on the CPU oracle: colours alternate from game to game, every
    move comes from a FRESH search of the side to move (select_action, :689-704: no root noise, temperature 0.12) with that side's
    net, the game ends on two passes or max_step, the winner is the area score against komi.  evaluate_match must play the same
    moves in every game (same stand-in nets, same uniform draws, first-maximum tie rule) and report the same winners."""
    from oracle import provider
    from oracle.go_oracle import OracleEnv
    from oracle.mcts_oracle import OracleMCTS, MCTSConfig
    from transgo_b200.selfplay import evaluate_match
    G, sims, max_step, seed = 6, 32, 14, 5
    seeds = {"new": (3, 6), "old": (4, 6)}
    out = evaluate_match(cpu_provider(*seeds["new"]), cpu_provider(*seeds["old"]), n_games=G, num_simulation=sims, max_step=max_step,
                         tie_mode="first", seed=seed, return_actions=True)
    acts = out["actions"].cpu().numpy()                                   # [plies, G]
    # the uniform draws of the batched match: per ply one vector per engine (new first, then old), slot k <-> k-th game of that side
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    half = G // 2
    new_black = np.arange(G) % 2 == 0
    idx_a, idx_b = np.nonzero(new_black)[0], np.nonzero(~new_black)[0]
    oenv = OracleEnv(9, max_step=max_step)
    states = [oenv.reset()[0] for _ in range(G)]
    done = [False] * G
    for ply in range(acts.shape[0]):
        black_to_move = ply % 2 == 0
        u_new = torch.rand([half], dtype=torch.float64, device="cuda", generator=gen).cpu().numpy()
        u_old = torch.rand([half], dtype=torch.float64, device="cuda", generator=gen).cpu().numpy()
        for who, idx, u in (("new", idx_a if black_to_move else idx_b, u_new), ("old", idx_b if black_to_move else idx_a, u_old)):
            sd, sharp = seeds[who]
            for k, g in enumerate(idx):
                if done[g]:
                    continue
                ora = OracleMCTS(oenv, lambda o, sd=sd, sharp=sharp: provider.synth_eval_batch(o, seed=sd, sharp=sharp),
                                 MCTSConfig(num_simulation=sims), u_fn=lambda m, u=u, k=k: float(u[k]))
                ora.root = ora._root_expand(states[g])                     # select_action: root = Node_V(state) + expand (:691-700)
                a, _, _ = ora.get_action_probs(is_selfplay=False)
                assert a == int(acts[ply, g]), (ply, g, a, int(acts[ply, g]))
                states[g], done[g] = oenv.step(states[g], a)
    assert all(done)
    winners = [1 if oenv.getScore(s) > 0 else 2 for s in states]          # environment.py:118-119
    assert winners == out["winners"].cpu().tolist()
    assert out["win"] == sum((w == 1) == nb for w, nb in zip(winners, new_black))


def test_continuous_selfplay_restarts_finished_slots(oracle_built):
    """continuous_self_play (self_play.py:902-968) batched: slots restart as their games end; every returned position belongs
    to a finished game, first positions are empty boards with black to move, labels are +-1 from the mover's view."""
    from transgo_b200.selfplay import SelfPlayWorker
    w = SelfPlayWorker(cpu_provider(7, 6), 5, num_simulation=16, max_step=6, tie_seed=1)
    out = w.play_continuous(min_positions=40, augment=False)
    n = out["positions"]
    assert n >= 40 and out["games"] >= 7 and out["obs"].shape == (n, 10, 9, 9) and out["pi"].shape == (n, 82)
    assert set(out["z"].cpu().tolist()) <= {1.0, -1.0} and bool(((out["pi"].sum(1) - 1).abs() < 1e-5).all())
    empties = int((out["obs"].reshape(n, -1).abs().sum(1) == 0).sum().item())
    assert empties == out["games"]                               # exactly one empty-board position per finished game
    assert n <= 6 * out["games"]                                 # max_step 6: at most 6 positions per game
    more = w.play_continuous(min_positions=10, augment=True)     # the unfinished games carry on
    assert more["obs"].shape[0] == 8 * more["positions"] and more["positions"] >= 10


# ------------------------------------------------------------------ parity at the BASELINE depths
def _hash_oracles(G, n, sims, seed, sharp, tie):
    from oracle import provider
    return make_oracles(G, n, sims, seed, sharp, tie=tie, eval_fn=lambda o: provider.hash_eval_batch(o, seed=seed, sharp=sharp))


@pytest.mark.parametrize("tie,graph", [("hash", True), ("first", False)])
def test_c3_depth_800_simulations_match_oracle(oracle_built, tie, graph):
    """BASELINE configs[2] depth: 800 simulations per move (deep trees, re-rooted arenas, 16-bit visit counters in use),
    3 games x 3 moves against the oracle; the evaluator is the device-side hash provider (bit-identical numpy twin on the
    oracle side), with the whole wave replayed as a CUDA graph in one of the two variants."""
    from helpers import torch_hash_provider
    from transgo_b200.mcts import BatchedMCTS
    seed, G, sims, sharp = 801, 3, 800, 5
    eng = BatchedMCTS(torch_hash_provider(seed, 9, sharp), G, num_simulation=sims, tie_mode=tie, tie_seed=seed, graph_waves=graph)
    compare_batch(eng, _hash_oracles(G, 9, sims, seed, sharp, tie), seed, 3)
    assert (eng._wave_graph is not None) == graph
    st = eng.search_stats()
    assert st["mean_depth"] > 2.0                                  # the trees really are deep


def test_c4_depth_19x19_400_simulations_match_oracle(oracle_built):
    """BASELINE configs[3] depth: 19x19, 400 simulations per move, 2 games x 2 moves, CUDA-graphed waves."""
    from helpers import torch_hash_provider
    from transgo_b200.mcts import BatchedMCTS
    seed, G, sims, sharp = 1904, 2, 400, 5
    eng = BatchedMCTS(torch_hash_provider(seed, 19, sharp), G, board_size=19, num_simulation=sims, tie_seed=seed, graph_waves=True)
    compare_batch(eng, _hash_oracles(G, 19, sims, seed, sharp, "hash"), seed, 2)


def test_select_action_from_a_mid_game_position_matches_oracle(oracle_built):
    """WP_MCTS.select_action(state) (self_play.py:689-704): a FRESH search from an arbitrary position, no root noise,
    temperature 0.12.  The position goes in as the reference-style state blob (set_root_state_blob); visit counts, pi and
    the root observation must equal the oracle's search from the same position."""
    from oracle import provider
    from oracle.go_oracle import OracleEnv
    from oracle.mcts_oracle import OracleMCTS, MCTSConfig
    from transgo_b200.environment import GoEnv
    from transgo_b200.mcts import WP_MCTS
    from helpers import playout_action
    seed, sims = 23, 210

    class Model:
        def main_prediction(self, x):
            p, v = provider.hash_eval_batch(x.cpu().numpy(), seed=seed, sharp=5)
            return torch.from_numpy(p).cuda(), torch.from_numpy(v).cuda(), torch.zeros(len(p), 81).cuda()

    class Cfg:
        board_size, num_simulation, c_puct1, c_puct2, wu_loss, parallel_readouts = 9, sims, 3, 0.05, 2, 4
    env, oenv = GoEnv(Cfg()), OracleEnv(9)
    agent = WP_MCTS(Cfg(), env, Model(), None, tie_mode="first")
    for plies in (37, 90):
        s, _ = env.reset()
        so, _ = oenv.reset()
        for t in range(plies):                                    # a mid-game position with captures and kos behind it
            pts = [x for x in oenv.legal_c(so) if x != 81]
            a = playout_action(5, plies, oenv.getStep(so), pts, 81)
            s, _ = env.step(s, a)
            so, _ = oenv.step(so, a)
        act = agent.select_action(s)
        ora = OracleMCTS(oenv, lambda o: provider.hash_eval_batch(o, seed=seed, sharp=5), MCTSConfig(num_simulation=sims))
        ora.root = ora._root_expand(so)                           # select_action: root = Node_V(state) + expand (:691-700)
        oa, opi, oobs = ora.get_action_probs(is_selfplay=False, forced_action=act)
        b = agent._b
        assert np.array_equal(b.counts[0].cpu().numpy(), ora.visit_counts()), plies
        assert int(b.root_n[0].item()) == ora.rootN
        counts = ora.visit_counts()
        assert counts[act] == counts.max() or counts[act] >= 2     # temperature 0.12 draw: a heavily visited move
        assert env.getStep(agent.root.state) == plies + 1 and env.checkAction(s, act if act != 81 else -1)


# ------------------------------------------------------------------ robustness
def test_capacity_limits_and_nan_fail_loudly(oracle_built):
    """edge_cap above the 22-bit edge-base field is refused (library and host side), a NaN from the evaluator raises a
    per-game error instead of walking a bogus edge, and max_depth defaults to the moves a game can still have."""
    from transgo_b200 import _lib
    from transgo_b200.mcts import BatchedMCTS
    ev = cpu_provider(1, 4)
    with pytest.raises(ValueError):
        BatchedMCTS(ev, 1, num_simulation=8, edge_cap=(1 << 22) + 1)
    eng = BatchedMCTS(ev, 1, num_simulation=8, edge_cap=1 << 22, node_cap=64)       # the boundary itself is fine
    assert eng.cfg.max_depth == 121 and eng.cfg.edge_cap == 1 << 22
    eng.cfg.edge_cap = (1 << 22) + 1
    assert eng.lib.tg_mcts_workspace_bytes(eng._cfg) < 0 and b"edge_cap" in eng.lib.tg_last_error()
    assert BatchedMCTS(ev, 1, board_size=19, num_simulation=2500).cfg.edge_cap <= 1 << 22      # the default caps clamp
    calls = [0]

    def nan_net(obs):
        p, v = ev(obs)
        calls[0] += 1
        if calls[0] >= 3:
            v = torch.full_like(v, float("nan"))
        return p, v
    bad = BatchedMCTS(nan_net, 2, num_simulation=40, tie_seed=3)
    with pytest.raises(_lib.TransgoLibraryError, match="NaN"):
        for _ in range(30):
            bad.run_wave()
        bad.check_status()


def test_continuous_selfplay_is_throttled_by_the_pacing_counters(oracle_built):
    """self_play.py:970-980: after a finished game the actor waits while now_train_steps / now_play_steps is below the
    train/play ratio.  play_continuous calls wait_fn (the trainer catching up) until the predicate clears, and returns
    early (throttled=True) when there is nobody to wait for."""
    from transgo_b200.replay import Pacing
    from transgo_b200.selfplay import SelfPlayWorker
    w = SelfPlayWorker(cpu_provider(7, 6), 4, num_simulation=12, max_step=6, tie_seed=1)
    pacing = Pacing()
    waits = [0]

    def trainer_catches_up():
        waits[0] += 1
        pacing.now_train_steps += 1
    out = w.play_continuous(min_positions=30, augment=False, pacing=pacing, wait_fn=trainer_catches_up)
    assert out["games"] == pacing.now_play_games >= 5 and pacing.now_play_steps == out["positions"]
    assert waits[0] >= 1 and not pacing.selfplay_should_wait() and out["throttled"] is False
    pacing.now_train_steps = 0
    out2 = w.play_continuous(min_positions=1000, augment=False, pacing=pacing)      # no trainer: hand control back
    assert out2["throttled"] is True and out2["positions"] < 1000


def test_pipelined_selection_equals_the_sequential_kernel():
    """k_select_par (one warp per descent of the wave, pipelined through shared memory, speculative with a sequential replay
    for terminal leaves / repeated ties) against k_select (one warp per game, the reference's descent order) at the BASELINE
    size: identical visit counts, actions, priors and statistics over several moves, including the end of short games."""
    from transgo_b200 import _lib
    from transgo_b200.mcts import BatchedMCTS, _declare
    L = _lib.lib(); _declare(L)
    for G, sims, max_step, moves in ((256, 200, None, 3), (64, 48, 10, 12)):
        outs = []
        for sequential in (1, 0):
            L.tg_mcts_debug_sequential_select(sequential)
            try:
                wave = [0]

                def evaluator(obs):
                    B = obs.shape[0]
                    p = torch.empty([B, 82], dtype=torch.float32, device="cuda"); v = torch.empty([B], dtype=torch.float32, device="cuda")
                    _lib.check(L.tg_synth_provider(9, B, 4242, wave[0], _lib.ptr(p), _lib.ptr(v), _lib.stream_ptr()))
                    wave[0] += 1
                    return p, v
                eng = BatchedMCTS(evaluator, G, num_simulation=sims, tie_seed=5, max_step=max_step)
                gen = torch.Generator(device="cuda"); gen.manual_seed(9)
                rec = []
                for _ in range(moves):
                    noise = eng.make_noise(generator=gen)
                    u = torch.rand([G], dtype=torch.float64, device="cuda", generator=gen)
                    a, pi, _ = eng.get_action_probs(True, noise=noise, u=u)
                    c, W, var, prior = eng.root_stats(full=True)
                    rec.append((a.clone(), pi.clone(), c.clone(), W.clone(), var.clone(), eng.root_n.clone()))
                    eng.update_with_action(a)
                outs.append((rec, eng.total_sims(), eng.search_stats()))
            finally:
                L.tg_mcts_debug_sequential_select(0)
        (ra, sa, sta), (rb, sb, stb) = outs
        assert sa == sb and sta == stb
        for x, y in zip(ra, rb):
            for p, q in zip(x, y):
                assert torch.equal(p, q)


def test_concurrent_shards_equal_one_engine(oracle_built):
    """ConcurrentMCTS (the batch cut into shards whose waves overlap on separate CUDA streams) must play exactly the games
    one engine plays: game ids -- and with them every hashed tie-break -- are global, games never interact."""
    from helpers import torch_hash_provider
    from transgo_b200.mcts import BatchedMCTS, ConcurrentMCTS
    seed, G, sims = 77, 16, 64
    one = BatchedMCTS(torch_hash_provider(seed, 9, 5), G, num_simulation=sims, tie_seed=seed, graph_waves=True)
    many = ConcurrentMCTS(torch_hash_provider(seed, 9, 5), G, shards=4, num_simulation=sims, tie_seed=seed, graph_waves=True)
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    for move in range(5):
        noise = one.make_noise(generator=gen)
        u = torch.rand([G], dtype=torch.float64, device="cuda", generator=gen)
        a1, p1, o1 = one.get_action_probs(True, noise=noise, u=u)
        a2, p2, o2 = many.get_action_probs(True, noise=noise, u=u)
        assert torch.equal(a1, a2) and torch.equal(p1, p2) and torch.equal(o1, o2), move
        assert torch.equal(one.counts, many.counts) and torch.equal(one.root_n, many.root_n)
        d1, d2 = one.update_with_action(a1), many.update_with_action(a2)
        assert torch.equal(d1, d2)
    assert one.total_sims() == many.total_sims()
    w1, s1, t1 = one.outcome(); w2, s2, t2 = many.outcome()
    assert torch.equal(w1, w2) and torch.equal(s1, s2) and torch.equal(t1, t2)
