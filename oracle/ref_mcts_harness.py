"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference search (self_play.py: WP_MCTS + GoEnv/environment.py +
go_env.so) under deterministic hooks.  Needs /root/reference, so it is used only in the build container: to generate
tests/golden/mcts_trace_*.npz and to cross-check oracle/mcts_oracle.py live.  Nothing is copied from the reference;
it is imported where it lies.

Hooks: `ray` is stubbed (not installed), np.random.dirichlet / np.random.choice are replaced while the
search runs so that noise, tie-breaks and move sampling are the injected ones, and the "model" is a stub whose
main_prediction is the deterministic provider.
"""
import contextlib
import os
import sys
import tempfile
import types

import numpy as np

REFERENCE = "/root/reference"


def available():
    return os.path.exists(os.path.join(REFERENCE, "self_play.py"))


def _import_reference():
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")

        def remote(*args, **kwargs):
            if len(args) == 1 and callable(args[0]) and not kwargs:
                return args[0]
            return lambda f: f
        ray.remote = remote
        ray.get = lambda x: x
        ray.init = lambda *a, **k: None
        sys.modules["ray"] = ray
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import self_play          # noqa: the reference module
    import configure
    return self_play, configure


class _StubModel:
    """Stands in for TransGoNetwork (model.py:11-27): main_prediction(obs[B,10,n,n]) -> (policy, value, own) tensors."""

    def __init__(self, eval_fn, n):
        import torch
        self.torch = torch
        self.eval_fn = eval_fn
        self.n = n
        self._p = torch.nn.Parameter(torch.zeros(1))

    def parameters(self):
        yield self._p

    def main_prediction(self, obs):
        torch = self.torch
        o = obs.detach().cpu().numpy()
        p, v = self.eval_fn(list(o))
        return torch.from_numpy(np.asarray(p, dtype=np.float32)), torch.from_numpy(np.asarray(v, dtype=np.float32)), \
            torch.zeros(len(o), self.n * self.n)


@contextlib.contextmanager
def _patched_random(state):
    old_d, old_c = np.random.dirichlet, np.random.choice

    def dirichlet(alpha, size=None):
        return np.asarray(state["noise_fn"](state["move_idx"], len(alpha)), dtype=np.float64)

    def choice(a, size=None, replace=True, p=None):
        if p is None:                                   # select_child tie-break, self_play.py:708-713
            a = list(a)
            return a[0] if len(a) == 1 else a[state["tie_fn"](len(a))]
        from .provider import choice_from_u
        if state.get("forced") is not None:
            return state["forced"]
        return int(np.asarray(a)[choice_from_u(p, state["u_fn"](state["move_idx"]))])
    np.random.dirichlet, np.random.choice = dirichlet, choice
    try:
        yield
    finally:
        np.random.dirichlet, np.random.choice = old_d, old_c


class ReferenceMCTS:
    """Same driving surface as oracle.mcts_oracle.OracleMCTS, backed by the reference's own WP_MCTS object."""

    def __init__(self, eval_fn, num_simulation=210, parallel_readouts=4, noise_fn=None, tie_fn=None, u_fn=None):
        import torch
        self_play, configure = _import_reference()
        cfg = configure.Config()
        cfg.device = torch.device("cpu")
        cfg.num_simulation = num_simulation
        cfg.parallel_readouts = parallel_readouts
        self.cfg = cfg
        cwd = os.getcwd()
        os.chdir(REFERENCE)                             # environment.py:42 loads ./GoEnv/go_env.so relative to CWD
        try:
            self.env = self_play.GoEnv(cfg)
        finally:
            os.chdir(cwd)
        self.state = dict(noise_fn=noise_fn, tie_fn=tie_fn or (lambda n: 0), u_fn=u_fn, move_idx=0, forced=None)
        with _patched_random(self.state), torch.no_grad():
            self.mcts = self_play.WP_MCTS(cfg, self.env, _StubModel(eval_fn, cfg.board_size), None)

    def get_action_probs(self, is_selfplay=True, forced_action=None):
        import torch
        self.state["forced"] = forced_action
        with _patched_random(self.state), torch.no_grad():
            a, pi, obs = self.mcts.get_action_probs(is_selfplay)
        return int(a), pi, obs

    def update_with_action(self, action):
        import torch
        with _patched_random(self.state), torch.no_grad():
            done = self.mcts.update_with_action(action)
        self.state["move_idx"] += 1
        return done

    @property
    def rootN(self):
        return self.mcts.root.total_visit_count

    def root_children(self):
        ch = self.mcts.root.children
        acts = np.array(sorted(ch.keys()), dtype=np.int64)
        N = np.array([ch[a].total_visit_count for a in acts], dtype=np.int64)
        W = np.array([ch[a].value_sum for a in acts], dtype=np.float32)
        var = np.array([ch[a].value_var for a in acts], dtype=np.float32)
        prior = np.array([ch[a].prior for a in acts], dtype=np.float64)
        return acts, N, W, var, prior
