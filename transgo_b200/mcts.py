"""Host-side mirror of the search entry point self_play.py calls (WP_MCTS, self_play.py:575-875).

``BatchedMCTS``  G concurrent searches on the GPU (one warp per game inside libtransgo_b200.so); same method names as
                 the reference with a leading game dimension: reset_root(), get_action_probs() -> (actions[G],
                 pi[G, A], obs[G, 10, n, n]), update_with_action(actions[G]) -> done[G].
``WP_MCTS``      the G = 1 facade with the reference's scalar signature (config, env, model, sub_model=None).

The policy/value net stays a PyTorch module (model.py: main_prediction); everything else -- selection, virtual loss,
leaf step / legal mask / feature planes, expansion, backup, re-rooting -- runs in hand-written CUDA.
"""
import ctypes
import math

import numpy as np

from . import _lib
from .environment import DEFAULT_MAX_STEP, GoState

ACTIVE, NEED_EXPAND, NOISY = 1, 2, 4
ERR_BITS = {16: "node arena overflow (raise node_cap)", 32: "edge arena overflow (raise edge_cap)",
            64: "search path deeper than max_depth", 128: "visit count above 65535",
            256: "update_with_action: action is not a child of the root",
            512: "no finite selection score at a node: the evaluator returned NaN / inf policy or value"}


class TgMctsConfig(ctypes.Structure):
    """include/transgo_b200.h: struct TgMctsConfig"""
    _fields_ = [("n", ctypes.c_int32), ("num_games", ctypes.c_int32), ("readouts", ctypes.c_int32),
                ("num_simulation", ctypes.c_int32), ("wu_loss", ctypes.c_int32), ("max_step", ctypes.c_int32),
                ("node_cap", ctypes.c_int32), ("edge_cap", ctypes.c_int32), ("max_depth", ctypes.c_int32),
                ("tie_mode", ctypes.c_int32), ("c_puct1", ctypes.c_double), ("c_puct2", ctypes.c_double),
                ("komi", ctypes.c_float), ("reserved0", ctypes.c_int32), ("tie_seed", ctypes.c_uint64),
                ("game_id0", ctypes.c_int64)]


def epsilon_by_frame(game_step):
    """configure.py:75-79 -- move-sampling temperature 1.0 -> 0.65"""
    return 0.65 + (1.0 - 0.65) * math.exp(-1.0 * game_step / 10)


def _declare(L):
    if getattr(L, "_tg_mcts_declared", False):
        return
    P = ctypes.POINTER(TgMctsConfig)
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    L.tg_mcts_workspace_bytes.argtypes = [P]
    L.tg_mcts_workspace_bytes.restype = ctypes.c_int64
    L.tg_mcts_init.argtypes = [P, vp, vp]
    L.tg_mcts_reset_games.argtypes = [P, vp, vp, vp]
    L.tg_mcts_root_obs.argtypes = [P, vp, vp, vp]
    L.tg_mcts_root_expand.argtypes = [P, vp, vp, vp, vp]
    L.tg_mcts_begin_move.argtypes = [P, vp, vp, vp]
    L.tg_mcts_wave_begin.argtypes = [P, vp, vp, vp]
    L.tg_mcts_wave_end.argtypes = [P, vp, vp, vp, vp, vp]
    L.tg_mcts_debug_sequential_select.argtypes = [i32]
    L.tg_mcts_pending.argtypes = [P, vp, vp, vp]
    L.tg_mcts_root_stats.argtypes = [P, vp] + [vp] * 7 + [vp]
    L.tg_mcts_advance.argtypes = [P, vp, vp, vp, vp]
    L.tg_mcts_status.argtypes = [P, vp, vp, vp, vp]
    L.tg_mcts_root_states.argtypes = [P, vp]
    L.tg_mcts_root_states.restype = ctypes.c_void_p
    L.tg_mcts_outcome.argtypes = [P, vp, vp, vp, vp, vp]
    L.tg_mcts_search_stats.argtypes = [P, vp, vp, vp]
    L.tg_synth_provider.argtypes = [i32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64, vp, vp, vp]
    L.tg_synth_provider_dev.argtypes = [i32, ctypes.c_int64, ctypes.c_uint64, vp, vp, vp, vp]
    L._tg_mcts_declared = True


class BatchedMCTS:
    """G searches in lock-step.  ``evaluator(obs) -> (policy [B, A] float32, value [B] or [B, 1] float32)`` is the net
    (obs is a CUDA float32 tensor [B, 10, n, n]; outputs must be CUDA tensors)."""

    def __init__(self, evaluator, num_games, board_size=9, num_simulation=210, c_puct1=3, c_puct2=0.05, wu_loss=2,
                 parallel_readouts=4, max_step=None, komi=7.5, tie_mode="hash", tie_seed=0, game_id0=0,
                 node_cap=None, edge_cap=None, max_depth=None, device="cuda", check_every=4, graph_waves=False):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise _lib.TransgoLibraryError("BatchedMCTS needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        _declare(self.lib)
        self.evaluator = evaluator
        self.G, self.n = int(num_games), int(board_size)
        self.P = self.n * self.n
        self.A = self.P + 1
        self.R = int(parallel_readouts)
        self.num_simulation = int(num_simulation)
        self.device = torch.device(device)
        self.check_every = check_every
        # graph_waves: capture one whole wave (select + leaves kernels, the evaluator's kernels, expand) as a CUDA graph on its
        # second call and replay it afterwards -- the wave loop is launch bound (a dozen launches around ~80 us of tree kernels).
        # The evaluator must then be capturable: device-only work on static tensors, no host synchronisation, no own graph.
        self.graph_waves = bool(graph_waves)
        self._wave_graph = None
        self._wave_calls = 0
        max_step = DEFAULT_MAX_STEP.get(self.n, 2 * self.P) if max_step is None else int(max_step)
        # every net-evaluated simulation adds one node; the kept subtree can carry several moves' worth of visits
        node_cap = node_cap or (6 * self.num_simulation + 64)
        # the per-edge child info is (edge base : 22 | count : 10): an arena holds at most 2^22 edges (the library checks too)
        edge_cap = edge_cap or min(1 << 22, max(self.A, int(node_cap * self.A * 0.85)))
        if edge_cap > (1 << 22):
            raise ValueError("edge_cap %d exceeds 4194304 (22-bit edge base in the per-edge child record)" % edge_cap)
        # the reference has no depth limit; a path can be as long as the moves left in the game
        max_depth = int(max_depth) if max_depth else max_step + 1
        self.cfg = TgMctsConfig(self.n, self.G, self.R, self.num_simulation, int(wu_loss), max_step, node_cap, edge_cap,
                                max_depth, 1 if tie_mode == "hash" else 0, float(c_puct1), float(c_puct2), float(komi), 0,
                                int(tie_seed), int(game_id0))
        self._cfg = ctypes.byref(self.cfg)
        nbytes = self.lib.tg_mcts_workspace_bytes(self._cfg)
        if nbytes <= 0:
            raise _lib.TransgoLibraryError("bad MCTS config: " + self.lib.tg_last_error().decode())
        self.workspace_bytes = int(nbytes)
        self.ws = torch.empty([(nbytes + 15) // 16, 4], dtype=torch.int32, device=self.device)
        G, R, A, n = self.G, self.R, self.A, self.n
        f32 = torch.float32
        self.obs = torch.zeros([G * R, 10, n, n], dtype=f32, device=self.device)        # the net's input, written in place
        self.obs_root = torch.zeros([G, 10, n, n], dtype=f32, device=self.device)
        self.pending = torch.zeros([1], dtype=torch.int32, device=self.device)
        self.counts = torch.zeros([G, A], dtype=torch.int32, device=self.device)
        self.root_n = torch.zeros([G], dtype=torch.int32, device=self.device)
        self.root_k = torch.zeros([G], dtype=torch.int32, device=self.device)
        self.step = torch.zeros([G], dtype=torch.int32, device=self.device)
        self.done = torch.zeros([G], dtype=torch.uint8, device=self.device)
        self.status = torch.zeros([G], dtype=torch.int32, device=self.device)
        self.sims = torch.zeros([1], dtype=torch.int64, device=self.device)
        self.waves = 0
        self._noise_seed, self._noise_calls = int(tie_seed) ^ 0x5DEECE66D, 0
        self.reset_root()

    # ---------------------------------------------------------------- plumbing
    def _s(self):
        return _lib.stream_ptr(device=self.device)

    def _w(self):
        return _lib.ptr(self.ws)

    def _eval(self, obs):
        out = self.evaluator(obs)
        policy, value = out[0], out[1]
        policy = policy.to(self.torch.float32).contiguous()
        value = value.to(self.torch.float32).reshape(-1).contiguous()
        assert policy.is_cuda and policy.shape == (obs.shape[0], self.A), policy.shape
        return policy, value

    def _expand_roots(self):
        """computeValuePolicy([root.state]) + root.expand(raw priors, value) for roots flagged NEED_EXPAND."""
        _lib.check(self.lib.tg_mcts_root_obs(self._cfg, self._w(), _lib.ptr(self.obs_root), self._s()), "tg_mcts_root_obs")
        policy, value = self._eval(self.obs_root)
        _lib.check(self.lib.tg_mcts_root_expand(self._cfg, self._w(), _lib.ptr(policy), _lib.ptr(value), self._s()),
                   "tg_mcts_root_expand")

    def check_status(self):
        _lib.check(self.lib.tg_mcts_status(self._cfg, self._w(), _lib.ptr(self.status), _lib.ptr(self.sims), self._s()),
                   "tg_mcts_status")
        st = self.status.cpu().numpy()
        bad = st & (16 | 32 | 64 | 128 | 256 | 512)
        if bad.any():
            g = int(np.nonzero(bad)[0][0])
            msgs = [m for b, m in ERR_BITS.items() if bad[g] & b]
            raise _lib.TransgoLibraryError("MCTS game %d: %s" % (g, "; ".join(msgs)))
        return st

    def total_sims(self):
        """Simulations (net-evaluated leaves + terminal backups) since the previous call."""
        self.check_status()
        n = int(self.sims.item())
        self.sims.zero_()
        return n

    def search_stats(self):
        """Selection statistics since the previous call: descents, mean depth (edges per descent), mean fan-out (children
        scanned per visited node) -- the d and K of the traffic model in SURVEY.md 8d."""
        out = self.torch.zeros([3], dtype=self.torch.int64, device=self.device)
        _lib.check(self.lib.tg_mcts_search_stats(self._cfg, self._w(), _lib.ptr(out), self._s()), "tg_mcts_search_stats")
        n, d, k = (int(x) for x in out.cpu().tolist())
        return dict(descents=n, mean_depth=d / max(1, n), mean_fanout=k / max(1, d))

    # ---------------------------------------------------------------- the reference surface
    def reset_root(self):                                   # self_play.py:595-605
        _lib.check(self.lib.tg_mcts_init(self._cfg, self._w(), self._s()), "tg_mcts_init")
        self._expand_roots()

    def reset_games(self, mask):
        """Start a fresh game in the slots where mask (uint8 [G]) is non-zero."""
        mask = mask.to(self.torch.uint8).contiguous()
        _lib.check(self.lib.tg_mcts_reset_games(self._cfg, self._w(), _lib.ptr(mask), self._s()), "tg_mcts_reset_games")
        self._expand_roots()

    def _wave_body(self):
        _lib.check(self.lib.tg_mcts_wave_begin(self._cfg, self._w(), _lib.ptr(self.obs), self._s()), "tg_mcts_wave_begin")
        policy, value = self._eval(self.obs)
        _lib.check(self.lib.tg_mcts_wave_end(self._cfg, self._w(), _lib.ptr(policy), _lib.ptr(value), None, self._s()),
                   "tg_mcts_wave_end")

    def games_pending(self):
        """Number of games that still need simulations for this move (one small launch + a host read)."""
        _lib.check(self.lib.tg_mcts_pending(self._cfg, self._w(), _lib.ptr(self.pending), self._s()), "tg_mcts_pending")
        return int(self.pending.item())

    def run_wave(self):                                     # self_play.py:607-654
        torch = self.torch
        if self._wave_graph is not None:
            self._wave_graph.replay()
        elif self.graph_waves and self._wave_calls >= 1:    # the first wave ran eagerly (lazy inits, autotuning); capture this one
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._wave_body()
            self._wave_graph = g
            g.replay()                                      # capture does not execute
        else:
            self._wave_body()
        self._wave_calls += 1
        self.waves += 1

    def make_noise(self, alpha=0.03, generator=None, seed=None):
        """Dirichlet(alpha) over each root's K children, float64 [G, A] in edge order (Node_V.dirichlet_prior :90-95), drawn on
        the device by tg_dirichlet (Philox counters keyed by seed / global game id / call number: shards draw what the full
        batch would).  `generator` (a torch.Generator) only supplies the seed."""
        torch = self.torch
        self.root_stats()
        if seed is None:
            if generator is not None:
                seed = int(torch.randint(0, 2 ** 62, [1], generator=generator, device=generator.device).item())
            else:
                seed = self._noise_seed
        self._noise_calls += 1
        noise = torch.empty([self.G, self.A], dtype=torch.float64, device=self.device)
        _lib.check(self.lib.tg_dirichlet(_lib.ptr(noise), _lib.ptr(self.root_k), self.G, self.A, float(alpha), int(seed) & (2 ** 64 - 1),
                                         self._noise_calls & 0xffffffff, int(self.cfg.game_id0), self._s()), "tg_dirichlet")
        return noise

    def root_stats(self, full=False):
        torch = self.torch
        W = var = prior = None
        if full:
            W = torch.empty([self.G, self.A], dtype=torch.float32, device=self.device)
            var = torch.empty_like(W)
            prior = torch.empty_like(W)
        _lib.check(self.lib.tg_mcts_root_stats(self._cfg, self._w(), _lib.ptr(self.counts), _lib.ptr(W), _lib.ptr(var),
                                               _lib.ptr(prior), _lib.ptr(self.root_n), _lib.ptr(self.root_k),
                                               _lib.ptr(self.step), self._s()), "tg_mcts_root_stats")
        return self.counts, W, var, prior

    def begin_move(self, noise=None):
        """Dirichlet mix at the roots (:90-95) and the simulation budget of this move (N0 = inherited visits, :660)."""
        _lib.check(self.lib.tg_mcts_begin_move(self._cfg, self._w(), _lib.ptr(noise), self._s()), "tg_mcts_begin_move")

    def search(self, noise=None):
        """The simulation loop of get_action_probs (:660-664) for all games."""
        self.begin_move(noise)
        min_waves = -(-self.num_simulation // self.R)
        for _ in range(min_waves):
            self.run_wave()
        while self.games_pending() > 0:
            for _ in range(self.check_every):
                self.run_wave()

    def get_action_probs(self, is_selfplay=True, noise=None, u=None, forced_actions=None):
        """Returns (actions int16 [G], pi float64 [G, A], root observations float32 [G, 10, n, n]).
        noise: float64 [G, A] Dirichlet draws in root-edge order (generated on the device when None and is_selfplay);
        u: float64 [G] uniforms for the move draw (torch.rand when None)."""
        noise = self.prepare_noise(is_selfplay, noise)
        self.search(noise)
        return self.pick_actions(is_selfplay, u, forced_actions)

    def prepare_noise(self, is_selfplay=True, noise=None):
        if not is_selfplay:
            return None
        if noise is None:
            noise = self.make_noise()
        return noise.to(self.torch.float64).contiguous()

    def pick_actions(self, is_selfplay=True, u=None, forced_actions=None):
        """The part of get_action_probs after the simulations (:666-687): pi from the visit counts, temperature draw, root obs."""
        torch = self.torch
        counts, _, _, _ = self.root_stats()
        c = counts.to(torch.float64)
        c1 = torch.where(counts == 1, torch.zeros_like(c), c)                   # :668
        # When no child was visited twice the reference divides by zero and np.random.choice raises (it needs ~2 x K
        # simulations to be safe; its 210 are).  Here such a game keeps its raw visit counts instead of failing the batch.
        c = torch.where(c1.sum(1, keepdim=True) > 0, c1, c)
        pi = c / c.sum(1, keepdim=True).clamp_min(1.0)                          # :670-671 (finished games: all zero)
        if forced_actions is not None:
            actions = forced_actions.to(torch.int16)
        else:
            if is_selfplay:                                                     # :676-677
                temp = 0.65 + 0.35 * torch.exp(-self.step.to(torch.float64) / 10.0)
            else:
                temp = torch.full([self.G], 0.12, dtype=torch.float64, device=self.device)
            tem = torch.pow(c, (1.0 / temp)[:, None])                            # :679
            cdf = torch.cumsum(tem / tem.sum(1, keepdim=True).clamp_min(1e-300), 1)
            cdf = cdf / cdf[:, -1:].clamp_min(1e-300)
            if u is None:
                u = torch.rand([self.G], dtype=torch.float64, device=self.device)
            actions = torch.searchsorted(cdf, u.to(torch.float64)[:, None].contiguous(), right=True).reshape(-1)
            actions = actions.clamp_max(self.A - 1).to(torch.int16)             # np.random.choice(p=...) :683
        _lib.check(self.lib.tg_mcts_root_obs(self._cfg, self._w(), _lib.ptr(self.obs_root), self._s()), "tg_mcts_root_obs")
        return actions, pi, self.obs_root                                       # :685-687

    def update_with_action(self, actions):                  # self_play.py:857-872
        actions = actions.to(self.torch.int16).contiguous()
        _lib.check(self.lib.tg_mcts_advance(self._cfg, self._w(), _lib.ptr(actions), _lib.ptr(self.done), self._s()),
                   "tg_mcts_advance")
        self._expand_roots()
        self.check_status()
        return self.done

    def outcome(self, want_territory=True):
        """winner uint8 [G], score float32 [G], territory float32 [G, P] of the current root states (self_play.py:931-937)."""
        torch = self.torch
        winner = torch.empty([self.G], dtype=torch.uint8, device=self.device)
        score = torch.empty([self.G], dtype=torch.float32, device=self.device)
        terr = torch.empty([self.G, self.P], dtype=torch.float32, device=self.device) if want_territory else None
        _lib.check(self.lib.tg_mcts_outcome(self._cfg, self._w(), _lib.ptr(winner), _lib.ptr(score), _lib.ptr(terr),
                                            self._s()), "tg_mcts_outcome")
        return winner, score, terr

    def root_state_blob(self, g=0):
        """Host copy of game g's root state in the single-state blob format of environment.GoEnv."""
        sq = self.lib.tg_state_bytes(self.n) // 16
        base = self.lib.tg_mcts_root_states(self._cfg, self._w())
        off = (base - self.ws.data_ptr()) // 4
        st = self.ws.reshape(-1)[off:off + sq * self.G * 4].reshape(sq, self.G, 4)[:, g, :].contiguous().cpu().numpy()
        blob = GoState()
        ctypes.memmove(ctypes.byref(blob), st.ctypes.data, sq * 16)
        return blob

    def set_root_state_blob(self, blob, g=0):
        torch = self.torch
        sq = self.lib.tg_state_bytes(self.n) // 16
        arr = np.frombuffer(bytes(blob.raw[:sq * 16]), dtype=np.int32).reshape(sq, 4)
        base = self.lib.tg_mcts_root_states(self._cfg, self._w())
        off = (base - self.ws.data_ptr()) // 4
        view = self.ws.reshape(-1)[off:off + sq * self.G * 4].reshape(sq, self.G, 4)
        view[:, g, :] = torch.from_numpy(arr.copy()).to(self.device)


class ConcurrentMCTS:
    """The same G searches run as `shards` independent BatchedMCTS engines on their own CUDA streams.

    Every tree kernel of a wave is latency bound at the BASELINE batch (one short dependent chain per game, a wave lasts as
    long as its slowest game), so one batch keeps the SMs mostly idle.  Games never interact: cutting the batch into shards
    whose waves are issued round-robin on separate streams lets the select of one shard overlap the leaves / evaluator /
    expand of the others.  Results are those of one engine over all games (game ids, hence every hashed draw, are global).

    `evaluators`: one callable per shard (each sees its own static [G/shards * R, 10, n, n] input tensor), or a single callable
    used by all shards.  Surface: the BatchedMCTS methods used by the self-play drivers."""

    def __init__(self, evaluators, num_games, shards=4, game_id0=0, device="cuda", graph_waves=False, waves_per_graph=1, **kw):
        import torch
        self.torch = torch
        assert num_games % shards == 0, "num_games must be divisible by shards"
        self.G, self.S = int(num_games), int(shards)
        # graph_waves: ONE CUDA graph holds `waves_per_graph` waves of every shard as parallel branches (fork / join inside the
        # capture), so a replay costs one launch on the host however many shards there are
        self.graph_waves, self.waves_per_graph = bool(graph_waves), int(waves_per_graph)
        self._graph, self._eager_rounds = None, 0
        per = self.G // self.S
        evs = list(evaluators) if isinstance(evaluators, (list, tuple)) else [evaluators] * self.S
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.S)]
        self.engines = []
        for i in range(self.S):
            with torch.cuda.stream(self.streams[i]):
                self.engines.append(BatchedMCTS(evs[i], per, game_id0=game_id0 + i * per, device=device, **kw))
        torch.cuda.synchronize()
        e0 = self.engines[0]
        self.n, self.A, self.P, self.R, self.num_simulation, self.cfg = e0.n, e0.A, e0.P, e0.R, e0.num_simulation, e0.cfg
        self.check_every = e0.check_every

    # -- stream plumbing: shard work is ordered after what the caller already queued, the caller after the shards
    def _fork(self):
        ev = self.torch.cuda.Event(); ev.record()
        for st in self.streams:
            st.wait_event(ev)

    def _join(self):
        cur = self.torch.cuda.current_stream()
        for st in self.streams:
            ev = self.torch.cuda.Event(); ev.record(st)
            cur.wait_event(ev)

    def _each(self, fn):
        self._fork()
        out = []
        for i, e in enumerate(self.engines):
            with self.torch.cuda.stream(self.streams[i]):
                out.append(fn(i, e))
        self._join()
        return out

    def _split(self, t):
        return [None] * self.S if t is None else list(t.chunk(self.S, 0))

    @property
    def waves(self):
        return sum(e.waves for e in self.engines)

    @property
    def workspace_bytes(self):
        return sum(e.workspace_bytes for e in self.engines)

    def reset_root(self):
        self._each(lambda i, e: e.reset_root())

    def reset_games(self, mask):
        ms = self._split(mask)
        self._each(lambda i, e: e.reset_games(ms[i]))

    def search(self, noise=None):
        torch = self.torch
        ns = self._split(noise)
        self._fork()
        for i, e in enumerate(self.engines):
            with torch.cuda.stream(self.streams[i]):
                e.begin_move(ns[i])
        min_waves = -(-self.num_simulation // self.R)
        if self.graph_waves:
            self._join()
            done = 0
            while done < min_waves or any(e.games_pending() > 0 for e in self.engines):
                done += self._round()
            return
        for _ in range(min_waves):                                    # waves issued round-robin: the shards overlap on the GPU
            for i, e in enumerate(self.engines):
                with torch.cuda.stream(self.streams[i]):
                    e.run_wave()
        live = list(range(self.S))
        while live:
            nxt = []
            for i in live:
                with torch.cuda.stream(self.streams[i]):
                    if self.engines[i].games_pending() > 0:
                        for _ in range(self.check_every):
                            self.engines[i].run_wave()
                        nxt.append(i)
            live = nxt
        self._join()

    def _round(self):
        """`waves_per_graph` waves of every shard; eager the first time (lazy initialisation), then captured once as a graph
        with one branch per shard and replayed.  Games whose budget is met sit the extra waves out (k_select skips them)."""
        torch = self.torch
        if self._graph is not None:
            self._graph.replay()
        elif self._eager_rounds >= 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._round_body()
            self._graph = g
            g.replay()
        else:
            self._round_body()
            self._eager_rounds += 1
        for e in self.engines:
            e.waves += self.waves_per_graph
        return self.waves_per_graph

    def _round_body(self):
        torch = self.torch
        cur = torch.cuda.current_stream()
        for i, e in enumerate(self.engines):
            st = self.streams[i]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for _ in range(self.waves_per_graph):
                    e._wave_body()
        for st in self.streams:
            cur.wait_stream(st)

    def get_action_probs(self, is_selfplay=True, noise=None, u=None, forced_actions=None):
        torch = self.torch
        ns = self._each(lambda i, e: e.prepare_noise(is_selfplay, self._split(noise)[i]))
        self.search(torch.cat(ns, 0) if is_selfplay else None)
        us, fs = self._split(u), self._split(forced_actions)
        outs = self._each(lambda i, e: e.pick_actions(is_selfplay, us[i], fs[i]))
        return tuple(torch.cat([o[k] for o in outs], 0) for k in range(3))

    def update_with_action(self, actions):
        acts = self._split(actions)
        return self.torch.cat(self._each(lambda i, e: e.update_with_action(acts[i])), 0)

    def root_stats(self, full=False):
        outs = self._each(lambda i, e: e.root_stats(full))
        return tuple(None if outs[0][k] is None else self.torch.cat([o[k] for o in outs], 0) for k in range(4))

    @property
    def counts(self):
        return self.torch.cat([e.counts for e in self.engines], 0)

    @property
    def root_n(self):
        return self.torch.cat([e.root_n for e in self.engines], 0)

    @property
    def step(self):
        return self.torch.cat([e.step for e in self.engines], 0)

    def outcome(self, want_territory=True):
        outs = self._each(lambda i, e: e.outcome(want_territory))
        return tuple(None if outs[0][k] is None else self.torch.cat([o[k] for o in outs], 0) for k in range(3))

    def check_status(self):
        return np.concatenate(self._each(lambda i, e: e.check_status()))

    def total_sims(self):
        return sum(self._each(lambda i, e: e.total_sims()))

    def search_stats(self):
        st = self._each(lambda i, e: e.search_stats())
        n = sum(x["descents"] for x in st)
        d = sum(x["descents"] * x["mean_depth"] for x in st)
        k = sum(x["descents"] * x["mean_depth"] * x["mean_fanout"] for x in st)
        return dict(descents=n, mean_depth=d / max(1, n), mean_fanout=k / max(1.0, d))


class _Root:
    def __init__(self, owner):
        self._o = owner

    @property
    def state(self):
        return self._o._b.root_state_blob(0)

    @property
    def total_visit_count(self):
        self._o._b.root_stats()
        return int(self._o._b.root_n[0].item())


class WP_MCTS(object):
    """Drop-in for self_play.py's WP_MCTS with the same constructor and entry points (self_play.py:906-926, :990-1016):
    reset_root(), get_action_probs(is_selfplay, now_train_step) -> (action, pi[A], obs[10, n, n]),
    update_with_action(a) -> done, select_action(state) -> action, .root.state, .model.  One game, run on the GPU."""

    def __init__(self, config, env, model, sub_model=None, **engine_kw):
        import torch
        self.config, self.env, self.model, self.sub_model = config, env, model, sub_model
        if sub_model is not None:
            raise NotImplementedError("the reference's transfer branch is dead code (self_play.py:809); not supported")
        self.board_size = config.board_size
        self._kw = dict(board_size=config.board_size, num_simulation=config.num_simulation, c_puct1=config.c_puct1,
                        c_puct2=config.c_puct2, wu_loss=config.wu_loss, parallel_readouts=config.parallel_readouts,
                        max_step=getattr(env, "max_step", None), komi=getattr(env, "komi", 7.5))
        self._kw.update(engine_kw)

        def evaluator(obs):
            with torch.no_grad():
                p, v, _ = self.model.main_prediction(obs)
            return p, v
        self._b = BatchedMCTS(evaluator, 1, **self._kw)      # constructor runs reset_root, like the reference (:593)
        self.root = _Root(self)

    def reset_root(self):
        self._b.reset_root()

    def get_action_probs(self, is_selfplay=True, now_train_step=0):
        a, pi, obs = self._b.get_action_probs(is_selfplay)
        return int(a[0].item()), pi[0].cpu().numpy(), obs[0].cpu().numpy().copy()

    def update_with_action(self, fall_action):
        torch = self._b.torch
        done = self._b.update_with_action(torch.tensor([int(fall_action)], dtype=torch.int16, device=self._b.device))
        return bool(done[0].item())

    def select_action(self, gamestate):                     # self_play.py:689-704
        b = self._b
        _lib.check(b.lib.tg_mcts_init(b._cfg, b._w(), b._s()), "tg_mcts_init")
        b.set_root_state_blob(gamestate, 0)
        b._expand_roots()
        action, _, _ = self.get_action_probs(is_selfplay=False)
        return action

    def __str__(self):
        return "WP_MCTS"
