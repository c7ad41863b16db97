"""CPU tests pinning the numpy MCTS restatement (oracle/mcts_oracle.py) to the reference search: it must reproduce the
golden trace the reference WP_MCTS produced, bit for bit (counts, W, var, pi, sampled action), and -- where
/root/reference is present -- match a live reference run."""
import numpy as np
import pytest

from helpers import load_mcts_golden, mcts_golden_games, golden_hooks, check_mcts_move_against_golden


def _oracle_engine(sims, sharp, seed, hooks):
    from oracle import provider
    from oracle.go_oracle import OracleEnv
    from oracle.mcts_oracle import OracleMCTS, MCTSConfig
    ev = lambda obs: provider.synth_eval_batch(obs, seed=seed, sharp=sharp)
    return OracleMCTS(OracleEnv(9), ev, MCTSConfig(num_simulation=sims), **hooks)


@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_mcts_oracle_reproduces_reference_golden(oracle_built, which):
    t = load_mcts_golden()
    game, sims, sharp, rows = list(mcts_golden_games(t))[which]
    rows = rows[:40]                                    # keeps the CPU suite short; the full games run in the GPU test
    hooks, ties = golden_hooks(t, rows, game)
    m = _oracle_engine(sims, sharp, int(t["seed"]), hooks)
    for i, row in enumerate(rows):
        a, pi, _ = m.get_action_probs()
        acts, N, W, var, _ = m.root_children()
        done = m.update_with_action(a)
        check_mcts_move_against_golden(t, row, acts, N, W, var, a, pi, done, rtol=0)
        assert ties[0] == int(t["nties"][row])
        if done:
            break


def test_mcts_oracle_matches_live_reference(oracle_built):
    from oracle import provider
    from oracle.ref_mcts_harness import ReferenceMCTS, available
    if not available():
        pytest.skip("/root/reference absent")
    seed, game, sims, sharp = 5, 9, 32, 7
    ev = lambda obs: provider.synth_eval_batch(obs, seed=seed, sharp=sharp)
    cnt = {"r": 0, "o": 0}

    def tie(which):
        def f(n):
            i = provider.tie_index(seed, game, cnt[which], n)
            cnt[which] += 1
            return i
        return f
    hooks = lambda w: dict(noise_fn=lambda m, k: provider.dirichlet(seed, game, m, k),
                           u_fn=lambda m: provider.uniform(seed, game, m), tie_fn=tie(w))
    ref = ReferenceMCTS(ev, num_simulation=sims, **hooks("r"))
    ora = _oracle_engine(sims, sharp, seed, hooks("o"))
    for m in range(30):
        a1, pi1, obs1 = ref.get_action_probs()
        a2, pi2, obs2 = ora.get_action_probs()
        r1, r2 = ref.root_children(), ora.root_children()
        for x, y in zip(r1[:4], r2[:4]):
            assert np.array_equal(x, y)
        assert a1 == a2 and np.array_equal(pi1, pi2) and np.array_equal(obs1, obs2)
        assert ref.update_with_action(a1) == ora.update_with_action(a2)
