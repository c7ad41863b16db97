"""TEST INFRASTRUCTURE ONLY -- ctypes faces of the two CPU checkers.

* ``OracleEnv``  : the plain-C restatement (oracle/go_rules.c -> liboracle_go.so).
* ``RefEnv``     : the UNMODIFIED reference engine compiled by oracle/Makefile into
                   oracle/_ref/go_env{9,19}.so (reference ABI: GoEnv/cpp_src/go_env.h:23-71).

Both expose the method names of the reference binding (GoEnv/environment.py:92-173) so parity
tests read like calls into the reference.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULTS = {9: dict(max_step=120, komi=7.5), 19: dict(max_step=722, komi=7.5)}


def build(verbose=False):
    """Compile the checkers (liboracle_go.so always; _ref/ only where /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def _load(path):
    if not os.path.exists(path):
        build()
    return ctypes.CDLL(path)


class _OgState(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("max_step", ctypes.c_int32), ("komi", ctypes.c_float),
                ("step_count", ctypes.c_int32), ("last1", ctypes.c_int32), ("last2", ctypes.c_int32),
                ("ko_loc", ctypes.c_int32), ("ko_age", ctypes.c_int32), ("next_player", ctypes.c_uint8),
                ("ko_color", ctypes.c_uint8), ("terminated", ctypes.c_uint8), ("color", ctypes.c_uint8 * 361)]


class OracleEnv:
    """Restated engine.  States are ``_OgState`` values (copy = fork, like the reference blobs)."""

    def __init__(self, board_size=9, max_step=None, komi=None):
        self.board_size = board_size
        self.P = board_size * board_size
        self.max_step = DEFAULTS[board_size]["max_step"] if max_step is None else max_step
        self.komi = DEFAULTS[board_size]["komi"] if komi is None else komi
        L = self.lib = _load(os.path.join(_HERE, "liboracle_go.so"))
        assert L.og_sizeof_state() == ctypes.sizeof(_OgState)
        S = ctypes.POINTER(_OgState)
        L.og_reset.argtypes = [S, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        L.og_step.argtypes = [S, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.og_step.restype = ctypes.c_int
        L.og_check_action.argtypes = [S, ctypes.c_int]
        L.og_legal.argtypes = [S, ctypes.c_void_p]
        L.og_legal_noeye.argtypes = [S, ctypes.c_void_p]
        L.og_encode10.argtypes = [S, ctypes.c_void_p]
        L.og_score.argtypes = [S]
        L.og_score.restype = ctypes.c_float
        L.og_territory.argtypes = [S, ctypes.c_void_p]
        L.og_territory.restype = ctypes.c_float
        L.og_sub_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4
        L.og_replay_playout.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong,
                                        ctypes.c_longlong, ctypes.c_uint64, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p]
        L.og_replay_playout.restype = ctypes.c_longlong
        L.og_batch_observe.argtypes = [ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_void_p] * 5
        L.og_batch_observe.restype = None

    def reset(self):
        s = _OgState()
        self.lib.og_reset(s, self.board_size, self.max_step, self.komi)
        return s, False

    def step(self, state, action):
        s = _OgState.from_buffer_copy(state)
        acc = ctypes.c_int(0)
        done = self.lib.og_step(s, int(action), ctypes.byref(acc))
        return s, bool(done)

    def checkAction(self, state, action):
        return bool(self.lib.og_check_action(state, int(action)))

    def encode(self, state):
        out = np.zeros([10, self.board_size, self.board_size], dtype=np.float32)
        self.lib.og_encode10(state, out.ctypes.data)
        return out

    def subEncode(self, enc, sub_board_size=7, cuts=4):
        enc = np.ascontiguousarray(enc, dtype=np.float32)
        out = np.zeros([cuts, enc.shape[0], sub_board_size, sub_board_size], dtype=np.float32)
        self.lib.og_sub_encode(enc.ctypes.data, out.ctypes.data, self.board_size, sub_board_size, enc.shape[0], cuts)
        return out

    def getScore(self, state):
        return float(self.lib.og_score(state))

    def getWinner(self, state):
        return 1 if self.getScore(state) > 0 else 2

    def legal_c(self, state):
        """The C-level list (PASS always appended), go_env.cc:154-171."""
        buf = np.zeros([self.P + 1], dtype=np.int32)
        n = self.lib.og_legal(state, buf.ctypes.data)
        return buf[:n]

    def getLegalAction(self, state):
        acts = self.legal_c(state)
        if len(acts) != 1:                      # environment.py:126-127: PASS only when alone
            acts = [a for a in acts if a != self.P]
        return acts

    def getLegalNoEye(self, state):
        buf = np.zeros([self.P + 1], dtype=np.int32)
        n = self.lib.og_legal_noeye(state, buf.ctypes.data)
        return buf[:n]

    def getScoreAndTerritory(self, state):
        t = np.zeros([self.P], dtype=np.float32)
        sc = self.lib.og_territory(state, t.ctypes.data)
        return float(sc), t

    # ---- whole-batch checkers (C loops; used to replay every game of a full-size playout launch)
    def replay_playout(self, trace, seed, game_id0=0, no_eye=False):
        """trace: int16 [plies, G] of device-chosen actions.  Verifies the move rule ply by ply (raises on the first
        mismatch) and returns (plies replayed, array of final states)."""
        trace = np.ascontiguousarray(trace, dtype=np.int16)
        plies, G = trace.shape
        finals = (_OgState * G)()
        bad = np.zeros(4, dtype=np.int32)
        n = self.lib.og_replay_playout(self.board_size, self.max_step, self.komi, trace.ctypes.data, plies, G, int(seed),
                                       int(game_id0), int(no_eye), ctypes.byref(finals), bad.ctypes.data)
        if n < 0:
            raise AssertionError("playout trace differs from the move rule: game %d ply %d expected %d got %d" % tuple(bad))
        return int(n), finals

    def batch_observe(self, finals):
        """(planes f32 [G,10,n,n], score f32 [G], territory f32 [G,P], legal bool [G,P+1] (python rule), noeye bool [G,P+1])"""
        G, n, P = len(finals), self.board_size, self.P
        planes = np.zeros([G, 10, n, n], np.float32); score = np.zeros([G], np.float32); terr = np.zeros([G, P], np.float32)
        legal = np.zeros([G, P + 1], np.uint8); noeye = np.zeros([G, P + 1], np.uint8)
        self.lib.og_batch_observe(ctypes.byref(finals), G, planes.ctypes.data, score.ctypes.data, terr.ctypes.data,
                                  legal.ctypes.data, noeye.ctypes.data)
        return planes, score, terr, legal.astype(bool), noeye.astype(bool)

    def getPlayer(self, state):
        return int(state.next_player)

    def getStep(self, state):
        return int(state.step_count)

    def isTerminated(self, state):
        return bool(state.terminated)


class RefEnv:
    """The compiled reference itself (oracle/_ref).  State = opaque blob, as in environment.py:17-29."""

    BLOB = 8192   # >= sizeof(State): 1188 B at 9x9, 5804 B at 19x19 with MAX_BLOCK=361

    def __init__(self, board_size=9, max_step=None, komi=None):
        self.board_size = board_size
        self.P = board_size * board_size
        self.max_step = DEFAULTS[board_size]["max_step"] if max_step is None else max_step
        self.komi = DEFAULTS[board_size]["komi"] if komi is None else komi
        path = os.path.join(_HERE, "_ref", "go_env%d.so" % board_size)
        L = self.lib = _load(path)
        L.Init.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]
        L.Init(1, 10, self.max_step, self.komi)          # environment.py:35-38,45
        for name in ("Reset", "isTerminated", "Show"):
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.Step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.Step.restype = ctypes.c_bool
        L.isTerminated.restype = ctypes.c_bool
        L.checkAction.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.checkAction.restype = ctypes.c_bool
        L.Encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.getScore.argtypes = [ctypes.c_void_p]
        L.getScore.restype = ctypes.c_float
        L.getTerritory.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.getTerritory.restype = ctypes.c_float
        L.getLegalAction.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.getLegalNoEye.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.getPlayer.argtypes = [ctypes.c_void_p]
        L.getPlayer.restype = ctypes.c_uint8
        L.getStep.argtypes = [ctypes.c_void_p]
        L.getSubEncode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]

    def _new(self):
        return ctypes.create_string_buffer(self.BLOB)

    def reset(self):
        s = self._new()
        self.lib.Reset(s)
        return s, False

    def step(self, state, action):
        s = self._new()
        done = self.lib.Step(state, s, int(action))
        return s, bool(done)

    def checkAction(self, state, action):
        return bool(self.lib.checkAction(state, int(action)))

    def encode(self, state):
        out = np.zeros([10, self.board_size, self.board_size], dtype=np.float32)
        self.lib.Encode(state, out.ctypes.data)
        return out

    def subEncode(self, enc, sub_board_size=7, cuts=4):
        enc = np.ascontiguousarray(enc, dtype=np.float32)
        out = np.zeros([cuts, enc.shape[0], sub_board_size, sub_board_size], dtype=np.float32)
        self.lib.getSubEncode(enc.ctypes.data, out.ctypes.data, sub_board_size, enc.shape[0], cuts)
        return out

    def getScore(self, state):
        return float(self.lib.getScore(state))

    def getWinner(self, state):
        return 1 if self.getScore(state) > 0 else 2

    def legal_c(self, state):
        buf = np.zeros([self.P + 1], dtype=np.int32)
        n = self.lib.getLegalAction(state, buf.ctypes.data)
        return buf[:n]

    def getLegalAction(self, state):
        acts = self.legal_c(state)
        if len(acts) != 1:
            acts = [a for a in acts if a != self.P]
        return acts

    def getLegalNoEye(self, state):
        buf = np.zeros([self.P + 1], dtype=np.int32)
        n = self.lib.getLegalNoEye(state, buf.ctypes.data)
        return buf[:n]

    def getScoreAndTerritory(self, state):
        t = np.zeros([self.P], dtype=np.float32)
        sc = self.lib.getTerritory(state, t.ctypes.data)
        return float(sc), t

    def getPlayer(self, state):
        return int(self.lib.getPlayer(state))

    def getStep(self, state):
        return int(self.lib.getStep(state))

    def isTerminated(self, state):
        return bool(self.lib.isTerminated(state))


def have_ref(board_size=9):
    return os.path.exists(os.path.join(_HERE, "_ref", "go_env%d.so" % board_size))
