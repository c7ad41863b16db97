"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the search the reference's self-play uses
(self_play.py: Node_V :51-96 and WP_MCTS :575-875), written edge-array style like the CUDA engine stores it.

It is the CHECKER for transgo_b200's MCTS kernels; nothing under transgo_b200/ imports it.  Parity status: PINNED --
tests/test_oracle_mcts.py requires it to reproduce tests/golden/mcts_trace_9.npz, which was produced by the reference
WP_MCTS itself (oracle/ref_mcts_harness.py, tests/golden/make_golden_mcts.py), and, where /root/reference is present,
to match a live run of the reference.

Numeric types follow the reference *as numpy 2.x executes it* (SURVEY.md 8a-3 "dtype ledger"):
  prior        float32 (policy[a] / scale, self_play.py:744); float64 at a root after dirichlet_prior (:95)
  W, mean, var float32 once touched; a never-visited child's var is the Python float 0.0  -> float64 var term
  ucb          (c1*prior [f32|f64]) * sqrt_f64(Np + ons_p) / (Nc + ons_c + 1)  +  var term  +  f32(-W/(Nc+1)), summed in f64
"""
import math

import numpy as np

f32 = np.float32
f64 = np.float64


class MCTSConfig:
    """The reference's search constants (configure.py:26-33, :75-79)."""

    def __init__(self, board_size=9, num_simulation=210, c_puct1=3, c_puct2=0.05, wu_loss=2, parallel_readouts=4):
        self.board_size = board_size
        self.num_simulation = num_simulation
        self.c_puct1 = c_puct1
        self.c_puct2 = c_puct2
        self.wu_loss = wu_loss
        self.parallel_readouts = parallel_readouts

    @staticmethod
    def temperature(game_step):                      # configure.py:75-79
        return 0.65 + (1.0 - 0.65) * math.exp(-1.0 * game_step / 10)


class Node:
    """An expanded node: its state and the statistics of its outgoing edges (= the reference's child Node_V objects)."""
    __slots__ = ("state", "acts", "prior", "noisy", "N", "ons", "W", "mean", "var", "child", "terminal_child")

    def __init__(self, state, acts, priors, w0):
        k = len(acts)
        self.state = state
        self.acts = np.asarray(acts, dtype=np.int64)
        self.prior = np.asarray(priors, dtype=f32)     # self_play.py:70-77 / :744
        self.noisy = None                              # float64 priors after dirichlet_prior
        self.N = np.zeros(k, dtype=np.int64)
        self.ons = np.zeros(k, dtype=np.int64)
        self.W = np.full(k, w0, dtype=f32)             # child.value_sum = -1 * value (self_play.py:76, :752)
        self.mean = np.zeros(k, dtype=f32)
        self.var = np.zeros(k, dtype=f32)
        self.child = [None] * k


class OracleMCTS:
    """One game.  `eval_fn(list_of_obs) -> (policy[B, A] f32, value[B, 1] f32)` stands in for model.main_prediction;
    hooks: noise_fn(move_idx, k) -> f64[k];  tie_fn(n) -> index;  u_fn(move_idx) -> uniform in [0, 1)."""

    def __init__(self, env, eval_fn, cfg, noise_fn=None, tie_fn=None, u_fn=None):
        self.env, self.eval_fn, self.cfg = env, eval_fn, cfg
        self.A = cfg.board_size ** 2 + 1
        self.noise_fn, self.tie_fn, self.u_fn = noise_fn, tie_fn or (lambda n: 0), u_fn
        self.move_idx = 0
        self.sims = 0          # net-evaluated leaves + terminal backups, for throughput accounting
        self.reset_root()

    # ------------------------------------------------------------------ roots
    def _root_expand(self, state):                     # self_play.py:595-605, :689-700, :862-870: RAW priors over legal
        policy, value = self.eval_fn([self.env.encode(state)])
        policy, value = policy[0], f32(value[0][0])
        legal = [int(a) for a in self.env.getLegalAction(state)]
        return Node(state, legal, [policy[a] for a in legal], f32(-1 * value))

    def reset_root(self):
        state, _ = self.env.reset()
        self.root = self._root_expand(state)
        self.rootN, self.root_ons = 0, 0
        self.rootW, self.root_mean, self.root_var = f32(0), f32(0), f32(0)
        self.move_idx = 0

    def dirichlet_prior(self):                         # self_play.py:90-95
        r = self.root
        noise = np.asarray(self.noise_fn(self.move_idx, len(r.acts)), dtype=f64)
        r.noisy = (r.prior * f32(1 - 0.25)).astype(f64) + noise * 0.25

    # ------------------------------------------------------------------ selection
    def _ucb(self, node, k, Np, onsp):                 # self_play.py:716-725
        c1, c2 = self.cfg.c_puct1, self.cfg.c_puct2
        sq = np.sqrt(Np + onsp)                        # np.sqrt(python int) -> float64
        if node.noisy is not None:
            ps = f64(c1) * node.noisy[k]
        else:
            ps = f64(f32(c1) * node.prior[k])
        ps = ps * sq / f64(int(node.N[k]) + int(node.ons[k]) + 1)
        if node.N[k] == 0:
            vs = f64(c2) * np.sqrt(f64(1.0))           # python-float var -> float64 path
        else:
            v = node.var[k]
            v = f32(0) if v < 0 else (f32(3) if v > 3 else v)     # np.clip(var, 0, 3) stays float32
            vs = f64(f32(c2) * np.sqrt(f32(1) + v))
        q = f64(-(node.W[k] / f32(int(node.N[k]) + 1)))          # -child.value(), float32 division
        return (ps + vs) + q

    def _select_child(self, node, Np, onsp):           # self_play.py:706-714
        scores = [self._ucb(node, k, Np, onsp) for k in range(len(node.acts))]
        best = max(scores)
        ties = [k for k, s in enumerate(scores) if s == best]
        return ties[0] if len(ties) == 1 else ties[self.tie_fn(len(ties))]

    # ------------------------------------------------------------------ statistics
    def _backup(self, path, value):                    # self_play.py:758-764 + Node_V.value_mean_var :84-88
        v = f32(value)
        for node, k in reversed(path):
            if node is None:
                self.rootW = f32(self.rootW + v); self.rootN += 1
                t = self.root_mean
                self.root_mean = f32(self.rootW / f32(self.rootN + 1))
                self.root_var = f32(self.root_var + f32(f32(v - t) * f32(v - self.root_mean)))
            else:
                node.W[k] = f32(node.W[k] + v); node.N[k] += 1
                t = node.mean[k]
                node.mean[k] = f32(node.W[k] / f32(int(node.N[k]) + 1))
                node.var[k] = f32(node.var[k] + f32(f32(v - t) * f32(v - node.mean[k])))
            v = f32(-v)

    def _add_ons(self, path, d):                       # self_play.py:767-774
        for node, k in path:
            if node is None:
                self.root_ons += d
            else:
                node.ons[k] += d

    # ------------------------------------------------------------------ one wave (self_play.py:607-654)
    def run(self):
        cfg, env = self.cfg, self.env
        paths, leaves = [], []
        failsafe = 0
        while len(paths) < cfg.parallel_readouts and failsafe < cfg.parallel_readouts * 2:
            failsafe += 1
            node, Np, onsp = self.root, self.rootN, self.root_ons
            path = [(None, -1)]                        # (owner node, edge index); (None, -1) is the root itself
            while True:
                k = self._select_child(node, Np, onsp)
                path.append((node, k))
                if node.child[k] is None:
                    break
                Np, onsp = int(node.N[k]), int(node.ons[k])
                node = node.child[k]
            parent, k = path[-1]
            leaf_state, done = env.step(parent.state, int(parent.acts[k]))
            if done:                                   # :638-642: scored at once, never expanded, not counted as a readout
                value = 1 if env.getPlayer(leaf_state) == env.getWinner(leaf_state) else -1
                self._backup(path, value)
                self.sims += 1
                continue
            self._add_ons(path, cfg.wu_loss)
            paths.append(path); leaves.append(leaf_state)
        if not paths:
            return
        policies, values = self.eval_fn([env.encode(s) for s in leaves])
        for path, leaf_state, policy, value in zip(paths, leaves, policies, values):
            self._add_ons(path, -cfg.wu_loss)
            parent, k = path[-1]
            if parent.child[k] is not None:            # :733-734: duplicate leaf of this wave -> no backup
                continue
            legal = [int(a) for a in env.getLegalAction(leaf_state)]
            scale = 0
            for a in legal:                            # python sum(): left-to-right float32 adds (:739)
                scale = scale + policy[a]
            v = f32(value[0])
            if scale > 0:
                node = Node(leaf_state, legal, [policy[a] / scale for a in legal], f32(-1 * v))
            else:                                      # children keep prior 0.0 and value_sum -0.0 (:634-636)
                node = Node(leaf_state, legal, [0.0] * len(legal), f32(-0.0))
            parent.child[k] = node
            self._backup(path, v)
            self.sims += 1

    # ------------------------------------------------------------------ per move (self_play.py:657-687, :857-872)
    def visit_counts(self):
        counts = np.zeros(self.A, dtype=np.int64)
        counts[self.root.acts] = self.root.N
        return counts

    def get_action_probs(self, is_selfplay=True, forced_action=None):
        if is_selfplay:
            self.dirichlet_prior()
        n0 = self.rootN
        while self.rootN < n0 + self.cfg.num_simulation:
            self.run()
        counts = self.visit_counts()
        counts = np.where(counts == 1, 0, counts)
        pi = counts / np.sum(counts)
        if is_selfplay:
            temperature = self.cfg.temperature(self.env.getStep(self.root.state))
        else:
            temperature = 0.12
        tem = np.power(counts, 1.0 / temperature)
        probs = np.array(tem) / np.sum(tem)
        if forced_action is not None:
            action = int(forced_action)
        else:
            from .provider import choice_from_u
            action = choice_from_u(probs, self.u_fn(self.move_idx))
        obs = self.env.encode(self.root.state)
        return action, pi, obs

    def update_with_action(self, action):
        r = self.root
        k = int(np.nonzero(r.acts == action)[0][0])
        next_state, done = self.env.step(r.state, action)
        self.rootN, self.root_ons = int(r.N[k]), int(r.ons[k])
        self.rootW, self.root_mean, self.root_var = r.W[k], r.mean[k], r.var[k]
        if r.child[k] is not None:
            self.root = r.child[k]
            self.root.noisy = None
        else:
            self.root = self._root_expand(next_state)
        self.move_idx += 1
        return done

    def root_children(self):
        """(acts, N, W, var, prior) of the root's edges -- what the parity tests compare."""
        r = self.root
        return r.acts.copy(), r.N.copy(), r.W.copy(), r.var.copy(), r.prior.copy()
