"""TEST INFRASTRUCTURE ONLY -- deterministic stand-ins for everything the reference leaves to chance or to the net.

The parity contract (BASELINE.json north_star): visit counts are bit-exact *given identical net outputs, Dirichlet
noise and tie-break order*.  These helpers produce those three inputs for every arm (the reference WP_MCTS run under
oracle/ref_mcts_harness.py, the numpy restatement oracle/mcts_oracle.py, and the CUDA engine):

* synth_eval     a "net": obs -> (policy f32[A], value f32), a pure function of the observation bytes built from
                 integer hashing and ONE IEEE float32 division per output, so it is bit-identical on any CPU and for
                 any batch composition (no exp/BLAS whose rounding depends on SIMD width).
* tie_index      which of n exactly-tied children select_child takes (stands in for np.random.choice, self_play.py:708-713)
* uniform / dirichlet   per-(game, move) draws from a Philox generator keyed by (seed, game, move)
"""
import zlib

import numpy as np

_M64 = (1 << 64) - 1


def mix32(seed, a, b):
    """Python-int twin of tg::mix32 (transgo_b200/csrc/tg_board.cuh) and tests/helpers.mix32."""
    z = (seed + a * 0x9E3779B97F4A7C15 + b * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 30; z = (z * 0xBF58476D1CE4E5B9) & _M64
    z ^= z >> 27; z = (z * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return z >> 32


def _mix32_vec(seed, a, b):
    with np.errstate(over="ignore"):
        z = (np.uint64(seed & _M64) + np.uint64(a & _M64) * np.uint64(0x9E3779B97F4A7C15)
             + np.asarray(b, dtype=np.uint64) * np.uint64(0xBF58476D1CE4E5B9))
        z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z >> np.uint64(32)


_CACHE = {}


def synth_eval(obs, seed=1234, sharp=3):
    """obs: float32 [10, n, n] (values 0/1).  Returns (policy float32 [A] summing to ~1, value float32 in (-1, 1))."""
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    A = obs.shape[-1] * obs.shape[-2] + 1
    h = zlib.crc32(np.packbits(obs.astype(np.uint8)).tobytes())
    key = (h, A, seed, sharp)
    hit = _CACHE.get(key)
    if hit is not None:
        return hit
    r = (_mix32_vec(seed, h, np.arange(A)) & np.uint64(0xFFFF)).astype(object)     # 16 bits, python ints from here
    w = np.array([1 + ((int(x) ** sharp) >> (16 * (sharp - 1))) for x in r], dtype=np.int64)   # heavy-tailed weights in [1, 65536]
    total = np.float32(int(w.sum()))
    policy = w.astype(np.float32) / total          # one correctly-rounded f32 division per entry
    v = (mix32(seed ^ 0x5555, h, 977) & 0xFFFF) - 32768
    value = np.float32(v) / np.float32(32768 + 64)
    if len(_CACHE) > 200000:
        _CACHE.clear()
    _CACHE[key] = (policy, value)
    return policy, value


def synth_eval_batch(obs_batch, seed=1234, sharp=3):
    ps, vs = zip(*[synth_eval(o, seed, sharp) for o in obs_batch])
    return np.stack(ps).astype(np.float32), np.array(vs, dtype=np.float32).reshape(-1, 1)


def tie_index(seed, game, counter, n):
    """Index (into the ascending-action list of tied children) that a tie-break takes; counter counts tie events."""
    return (mix32(seed, game, counter) * n) >> 32


def _gen(seed, game, move, salt):
    return np.random.Generator(np.random.Philox(key=[(seed * 1000003 + salt) & _M64, (game << 20) + move]))


def dirichlet(seed, game, move, k, alpha=0.03):
    """Root noise of Node_V.dirichlet_prior (self_play.py:90-95): float64 [k]."""
    return _gen(seed, game, move, 1).dirichlet([alpha] * k)


def uniform(seed, game, move):
    """The uniform draw np.random.choice(p=...) consumes in get_action_probs (self_play.py:683)."""
    return float(_gen(seed, game, move, 2).random())


def choice_from_u(p, u):
    """numpy's Generator/RandomState.choice(p=...) algorithm: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(u, 'right')."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


# ---- a second stand-in net whose arithmetic is also expressible as exact int64 tensor ops + one float32 division, so the
# CUDA engine can evaluate it ON THE DEVICE (tests/helpers.py: torch_hash_provider; capturable in a CUDA graph) while the
# oracle evaluates this numpy twin: identical bits on both sides.
def hash_eval_tables(seed, n_inputs, A):
    coef = (_mix32_vec(seed, 11, np.arange(n_inputs)) | np.uint64(1)).astype(np.int64)        # < 2^32
    mul = ((_mix32_vec(seed, 12, np.arange(A)) >> np.uint64(1)) | np.uint64(1)).astype(np.int64)   # < 2^31
    add = _mix32_vec(seed, 13, np.arange(A)).astype(np.int64)
    return coef, mul, add


def hash_eval_batch(obs_batch, seed=77, sharp=3):
    """obs [B, 10, n, n] of 0/1 -> (policy f32 [B, A], value f32 [B, 1]); heavy-tailed integer weights / their sum."""
    obs = np.asarray(obs_batch, dtype=np.float32)
    B = obs.shape[0]
    A = obs.shape[-1] * obs.shape[-2] + 1
    bits = obs.reshape(B, -1).astype(np.int64)
    coef, mul, add = hash_eval_tables(seed, bits.shape[1], A)
    h = (bits * coef[None, :]).sum(1) & np.int64(0x7FFFFFFF)
    r = ((h[:, None] * mul[None, :] + add[None, :]) >> np.int64(13)) & np.int64(0xFFFF)
    w = r
    for _ in range(sharp - 1):
        w = (w * r) >> np.int64(16)
    w = w + np.int64(1)
    total = w.sum(1).astype(np.float32)
    policy = w.astype(np.float32) / total[:, None]
    v = (((h * np.int64(2654435761)) >> np.int64(9)) & np.int64(0xFFFF)) - np.int64(32768)
    value = v.astype(np.float32) / np.float32(32832.0)
    return policy.astype(np.float32), value.reshape(-1, 1).astype(np.float32)
