"""Host-side mirror of the reference environment binding (GoEnv/environment.py:32-198).

``GoEnv``         same constructor, method names, argument meaning, return conventions and error behaviour as
                  the reference class, for ONE state at a time; every rules call runs the sm_100a kernels
                  through the reference-named C symbols of libtransgo_b200.so.
``BatchedGoEnv``  the B200 hot path: G games resident in HBM, one kernel launch per call, torch tensors in/out.
"""
import ctypes

import numpy as np

from . import _lib

BLACK = 1     # environment.py:14-15
WHITE = 2

DEFAULT_MAX_STEP = {9: 120, 19: 722}   # environment.py:37; 19x19 has no reference default (SURVEY.md 8d C4 note)


class GoState(ctypes.Structure):
    """Opaque state value.  The reference blob (environment.py:17-29) is 1196 bytes and never field-accessed;
    this one is padded to the same size at 9x9 so either library can be handed either buffer."""
    _fields_ = [("raw", ctypes.c_uint8 * 1196)]


class GoEnv:
    def __init__(self, config=None, board_size=None):
        self.config = config
        self.history_dim = 1            # environment.py:35-40
        self.encoded_dim = 10
        self.board_size = board_size or (getattr(config, "board_size", 9) if config is not None else 9)
        self.max_step = DEFAULT_MAX_STEP.get(self.board_size, 2 * self.board_size ** 2)
        self.komi = 7.5
        self.sub_board_size = 7
        self.lib = _lib.lib()
        if not self.lib.tg_set_board_size(self.board_size):
            raise ValueError("unsupported board size %d" % self.board_size)
        self.lib.Init(self.history_dim, self.encoded_dim, self.max_step, self.komi)
        self._P = self.board_size ** 2

    # -- environment.py:92-103
    def reset(self):
        new_state = GoState()
        self.lib.Reset(ctypes.byref(new_state))
        return new_state, False

    def step(self, state, action):
        new_state = GoState()
        done = self.lib.Step(ctypes.byref(state), ctypes.byref(new_state), int(action))
        return new_state, done

    # -- environment.py:105-113
    def encode(self, state):
        out = np.zeros([self.history_dim * self.encoded_dim, self.board_size, self.board_size], dtype="float32")
        self.lib.Encode(ctypes.byref(state), out.ctypes.data)
        return out

    def subEncode(self, encode):
        encode = np.ascontiguousarray(encode, dtype="float32")
        sub = np.zeros([4, self.encoded_dim, self.sub_board_size, self.sub_board_size], dtype="float32")
        self.lib.getSubEncode(encode.ctypes.data, sub.ctypes.data, self.sub_board_size, self.encoded_dim, 4)
        return sub

    # -- environment.py:115-129
    def getScore(self, state):
        return self.lib.getScore(ctypes.byref(state))

    def getWinner(self, state):
        return BLACK if self.getScore(state) > 0 else WHITE

    def getLegalAction(self, state):
        legal_action = np.zeros([self._P + 1], dtype="int32")
        num_action = self.lib.getLegalAction(ctypes.byref(state), legal_action.ctypes.data)
        legal_acts = legal_action[:num_action]
        if num_action != 1:
            legal_acts = [act for act in legal_acts if act != self._P]
        return legal_acts

    def getPlayer(self, state):
        return self.lib.getPlayer(ctypes.byref(state))

    def action_to_location(self, action):
        return [action // self.board_size, action % self.board_size]

    def location_to_action(self, location):
        return self.board_size * location[0] + location[1]

    # -- environment.py:153-173
    def checkAction(self, state, action):
        return self.lib.checkAction(ctypes.byref(state), int(action))

    def getScoreAndTerritory(self, state):
        territory = np.zeros([self._P], dtype="float32")
        score = self.lib.getTerritory(ctypes.byref(state), territory.ctypes.data)
        return score, territory

    def getLegalNoEye(self, state):
        candidate = np.zeros([self._P + 1], dtype="int32")
        n = self.lib.getLegalNoEye(ctypes.byref(state), candidate.ctypes.data)
        return candidate[:n]

    def show(self, state):
        self.lib.Show(ctypes.byref(state))

    def getStep(self, state):
        return self.lib.getStep(ctypes.byref(state))

    def justStarted(self, state):
        return self.lib.getStep(ctypes.byref(state)) == 1

    def isTerminated(self, state):
        return self.lib.isTerminated(ctypes.byref(state))


class BatchedGoEnv:
    """G concurrent games in HBM.  State tensor: int32 [SQ, G, 4] (SQ planes of 128-bit words, include/transgo_b200.h).

    ``step`` keeps the reference semantics per game (go_env.cc:51-80): terminated games stay put and report
    done, illegal actions leave the state unchanged and report accepted = 0.

    ``tile`` = lanes cooperating on one game: 0 lets the library choose from (board, G) by measurement (see
    ``playout_layout``); 32 / 16 / 8 / 4 / 2 / 1 force a bitboard tile; ``n | 64`` (n in 16, 8, 4) runs the tiles that share
    a warp in lock step; ``n | 128`` (n in 32, 16, 8, 4, 1) is the incremental engine (dense bit-string boards, atari set kept
    across plies: csrc/tg_inc.cuh -- what the library picks); -1 is the row-per-lane layout.  The last three only exist for
    ``playout``; the other kernels then fall back to the library's choice.  Every layout produces identical results
    (tests/test_env_gpu.py).
    """

    def __init__(self, num_games, board_size=9, max_step=None, komi=7.5, device="cuda", tile=0, step_engine="incremental"):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise _lib.TransgoLibraryError("BatchedGoEnv needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        if not self.lib.tg_supported_board(board_size):
            raise ValueError("unsupported board size %d" % board_size)
        self.G = int(num_games)
        self.n = board_size
        self.P = board_size * board_size
        self.A = self.P + 1
        self.max_step = DEFAULT_MAX_STEP.get(board_size, 2 * self.P) if max_step is None else int(max_step)
        self.komi = float(komi)
        self.device = torch.device(device)
        self.tile = tile
        self.SQ = self.lib.tg_state_bytes(board_size) // 16
        self.LW = self.lib.tg_legal_words(board_size)
        self.states = torch.zeros([self.SQ, self.G, 4], dtype=torch.int32, device=self.device)
        self.legal = torch.zeros([self.G, self.LW], dtype=torch.int32, device=self.device)
        self.done = torch.zeros([self.G], dtype=torch.uint8, device=self.device)
        self.accepted = torch.zeros([self.G], dtype=torch.uint8, device=self.device)
        # step_engine "incremental" (default): tg_env_step_inc, the atari sets kept beside the positions in `self.atari` (a cache:
        # records that do not match their position are rebuilt inside the launch, so writing to `states` directly stays safe);
        # "analysis": tg_env_step, which re-derives what it needs on every call (the cross-check, and `tile` applies to it only)
        assert step_engine in ("incremental", "analysis")
        self.step_engine = step_engine
        self.atari = torch.zeros([self.G, self.lib.tg_env_atari_words(board_size)], dtype=torch.int32, device=self.device)
        self._bitpos = None
        self.reset()

    def _s(self):
        return _lib.stream_ptr(device=self.device)

    def reset(self):
        _lib.check(self.lib.tg_env_reset(self.n, _lib.ptr(self.states), self.G, self._s()), "tg_env_reset")
        self.done.zero_()
        return self.states

    def step(self, actions, want_legal=True):
        """actions: int16 [G] on the device.  Returns (legal mask int32 [G, LW] of the new positions, done uint8 [G])."""
        assert actions.dtype == self.torch.int16 and actions.is_cuda and actions.numel() == self.G
        if self.step_engine == "incremental":
            _lib.check(self.lib.tg_env_step_inc(self.n, _lib.ptr(self.states), _lib.ptr(self.atari), self.G, _lib.ptr(actions),
                                                _lib.ptr(self.legal) if want_legal else None, _lib.ptr(self.done),
                                                _lib.ptr(self.accepted), self.max_step, self._s()), "tg_env_step_inc")
            return self.legal, self.done
        _lib.check(self.lib.tg_env_step(self.n, _lib.ptr(self.states), self.G, _lib.ptr(actions),
                                        _lib.ptr(self.legal) if want_legal else None, _lib.ptr(self.done),
                                        _lib.ptr(self.accepted), self.max_step, self.tile, self._s()), "tg_env_step")
        return self.legal, self.done

    def make_step_graph(self, planes=None):
        """Capture `step` (+ `encode` into ``planes`` when given) once as a CUDA graph and return ``replay(actions) -> (legal, done)``.
        A ply-by-ply driver pays one graph launch per ply instead of one C-ABI call + kernel launch per kernel (the kernels of a ply
        last a few microseconds at the BASELINE batch: the host side is what a stepwise loop waits for).  ``actions`` are copied
        into a static buffer; ``legal`` / ``done`` / ``planes`` are the env's own tensors, overwritten by every replay."""
        torch = self.torch
        static_actions = torch.zeros([self.G], dtype=torch.int16, device=self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        keep = self.states.clone()
        with torch.cuda.stream(side):                          # warm-up outside the capture (lazy initialisation)
            self.step(static_actions)
            if planes is not None:
                self.encode(out=planes)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.states.copy_(keep)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.step(static_actions)
            if planes is not None:
                self.encode(out=planes)
        self.states.copy_(keep)                                # the capture does not execute, but stay explicit

        def replay(actions):
            static_actions.copy_(actions, non_blocking=True)
            graph.replay()
            return self.legal, self.done
        replay.graph = graph
        return replay

    def legal_mask(self, no_eye=False, out=None):
        out = self.legal if out is None else out
        _lib.check(self.lib.tg_env_legal(self.n, _lib.ptr(self.states), self.G, _lib.ptr(out), int(no_eye), self.tile,
                                         self._s()), "tg_env_legal")
        return out

    def encode(self, out=None):
        """float32 [G, 10, n, n]; pass ``out`` (e.g. a slice of the net's input tensor) to write in place."""
        if out is None:
            out = self.torch.empty([self.G, 10, self.n, self.n], dtype=self.torch.float32, device=self.device)
        assert out.is_contiguous() and out.dtype == self.torch.float32
        _lib.check(self.lib.tg_env_encode(self.n, _lib.ptr(self.states), self.G, _lib.ptr(out), self.tile, self._s()),
                   "tg_env_encode")
        return out

    def score(self, want_territory=False):
        sc = self.torch.empty([self.G], dtype=self.torch.float32, device=self.device)
        terr = self.torch.empty([self.G, self.P], dtype=self.torch.float32, device=self.device) if want_territory else None
        _lib.check(self.lib.tg_env_score(self.n, _lib.ptr(self.states), self.G, self.komi, _lib.ptr(sc), _lib.ptr(terr),
                                         self.tile, self._s()), "tg_env_score")
        return (sc, terr) if want_territory else sc

    def playout(self, plies, seed=0x7261, game_id0=0, no_eye=False, trace=None, counter=None, from_reset=False, score_out=None,
                states=None, states_in=None):
        """Every game plays up to ``plies`` moves of the SURVEY 8d C2 rule in one launch.  ``from_reset`` starts from empty boards
        (no separate reset launch, the old states are not read); ``score_out`` (float32 [G]) receives the final area scores
        (no separate score launch).  ``states`` (int32 [SQ, G, 4]) plays that buffer in place instead of ``self.states``.
        ``states``, ``trace`` and ``score_out`` may be PINNED HOST tensors: the kernel then reads the positions and writes the
        moves / results across PCIe itself (zero-copy; a board lives in registers for the whole launch, so each buffer
        crosses the bus once) -- for a host caller that is one launch instead of a launch and four copies.  ``states_in`` (same
        shape) is read instead of ``states``, which is then only written: start positions kept in a read-only buffer need no
        refill between launches."""
        st = self.states if states is None else states
        assert st.dtype == self.torch.int32 and st.numel() == self.SQ * self.G * 4 and st.is_contiguous()
        assert st.is_cuda or st.is_pinned(), "states must live on the device or in pinned host memory"
        for t in (trace, score_out):
            assert t is None or t.is_cuda or t.is_pinned(), "trace / score_out must live on the device or in pinned host memory"
        if states_in is not None:
            assert not from_reset and states_in.dtype == self.torch.int32 and states_in.numel() == st.numel() and states_in.is_contiguous()
            assert states_in.is_cuda or states_in.is_pinned(), "states_in must live on the device or in pinned host memory"
            _lib.check(self.lib.tg_env_playout_io(self.n, _lib.ptr(states_in), _lib.ptr(st), self.G, int(plies), self.max_step, int(seed),
                                                  int(game_id0), int(no_eye), _lib.ptr(trace), _lib.ptr(counter), _lib.ptr(score_out),
                                                  self.komi, self.tile, self._s()), "tg_env_playout_io")
            return
        if from_reset or score_out is not None:
            _lib.check(self.lib.tg_env_playout_scored(self.n, _lib.ptr(st), self.G, int(plies), self.max_step, int(seed),
                                                      int(game_id0), int(no_eye), _lib.ptr(trace), _lib.ptr(counter), int(from_reset),
                                                      _lib.ptr(score_out), self.komi, self.tile, self._s()), "tg_env_playout_scored")
            if from_reset:
                self.done.zero_()
            return
        _lib.check(self.lib.tg_env_playout(self.n, _lib.ptr(st), self.G, int(plies), self.max_step, int(seed),
                                           int(game_id0), int(no_eye), _lib.ptr(trace), _lib.ptr(counter), self.tile,
                                           self._s()), "tg_env_playout")

    def playout_layout(self):
        """Name of the kernel ``playout`` launches for this batch: lanes per game, or 'rows' (one board row per lane)."""
        t = self.lib.tg_env_playout_tile(self.n, self.G, self.tile)
        if t < 0:
            return "rows"
        if t & 128:
            w = t & 63
            return "incremental, " + ("one thread per game" if w == 1 else "%d lanes per game, %d games per warp in lock step" % (w, 32 // w))
        return "%d lanes per game, %d games per warp in lock step" % (t & 63, 32 // (t & 63)) if t & 64 else str(t)

    # ---- helpers (format conversion only)
    def unpack_legal(self, legal=None):
        """dense mask int32 [G, LW] -> bool [G, A]"""
        torch = self.torch
        legal = self.legal if legal is None else legal
        if self._bitpos is None:
            a = torch.arange(self.A, device=self.device)
            self._bitpos = (a // 32, (a % 32).to(torch.int32))
        w, b = self._bitpos
        return ((legal[:, w] >> b) & 1).bool()

    def meta(self):
        """Decoded meta words as a dict of int tensors (step, player, last1, last2, ko_loc, ko_age, terminated)."""
        torch = self.torch
        m = self.states[self.SQ - 1]          # [G, 4]
        s16 = lambda v: ((v & 0xffff) ^ 0x8000) - 0x8000
        return dict(last1=s16(m[:, 0]), last2=s16(m[:, 0] >> 16), ko_loc=s16(m[:, 1]), ko_age=s16(m[:, 1] >> 16),
                    step=m[:, 2] & 0xffff, player=(m[:, 2] >> 16) & 0xff, ko_color=(m[:, 2] >> 24) & 0xff,
                    terminated=m[:, 3] & 1)
