"""dev: per-game duration of the pipelined selection launch at the C3 size (globaltimer stamps written by the kernel)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
from transgo_b200.mcts import BatchedMCTS, _declare
L = _lib.lib(); _declare(L)
L.tg_mcts_debug_times.argtypes = [ctypes.c_void_p]
G = 1024
pol = torch.empty([G * 4, 82], dtype=torch.float32, device="cuda"); val = torch.empty([G * 4], dtype=torch.float32, device="cuda")
pol_r = torch.empty([G, 82], dtype=torch.float32, device="cuda"); val_r = torch.empty([G], dtype=torch.float32, device="cuda")
wave = torch.zeros([2], dtype=torch.int64, device="cuda")


def synth(obs):
    p, v = (pol, val) if obs.shape[0] == G * 4 else (pol_r, val_r)
    _lib.check(L.tg_synth_provider_dev(9, obs.shape[0], 5, _lib.ptr(wave), _lib.ptr(p), _lib.ptr(v), _lib.stream_ptr()))
    return p, v


eng = BatchedMCTS(synth, G, num_simulation=800, tie_seed=1)
a, _, _ = eng.get_action_probs(True); eng.update_with_action(a)
L.tg_mcts_begin_move(eng._cfg, eng._w(), None, eng._s())
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    eng.run_wave()
times = torch.zeros([G, 4], dtype=torch.int64, device="cuda")
_lib.check(L.tg_mcts_debug_times(_lib.ptr(times)))
for rep in range(3):
    eng.run_wave()
    torch.cuda.synchronize()
    t = times.cpu().numpy().astype("int64")
    t0 = t[:, 0].min()
    start, sel, end, depth = t[:, 0] - t0, t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3]
    import numpy as np
    q = lambda a: "min %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(np.percentile(a, [0, 50, 90, 99, 100]) / 1e3)
    print("wave %d: kernel span %.1f us | CTA start skew: %s | selection: %s | leaves: %s | CTA total: %s" % (
        rep, (t[:, 2].max() - t0) / 1e3, q(start), q(sel), q(end), q(t[:, 2] - t[:, 0])))
    print("   deepest path per game: p50 %d p90 %d p99 %d max %d | selection us by depth: %s" % (
        *np.percentile(depth, [50, 90, 99, 100]).astype(int), ", ".join("%d:%.1f" % (d, sel[depth == d].mean() / 1e3) for d in sorted(set(depth.tolist()))[:14])))
_lib.check(L.tg_mcts_debug_times(None))
