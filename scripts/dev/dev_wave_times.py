"""Per-kernel device time of one MCTS wave (CUDA events around each C-ABI call, eager launches, 1024 games x 4 leaves)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
from transgo_b200.mcts import BatchedMCTS, _declare
L = _lib.lib(); _declare(L)
G = 1024
pol = torch.empty([G * 4, 82], dtype=torch.float32, device="cuda"); val = torch.empty([G * 4], dtype=torch.float32, device="cuda")
pol_r = torch.empty([G, 82], dtype=torch.float32, device="cuda"); val_r = torch.empty([G], dtype=torch.float32, device="cuda")
wave = torch.zeros([2], dtype=torch.int64, device="cuda")


def synth(obs):
    p, v = (pol, val) if obs.shape[0] == G * 4 else (pol_r, val_r)
    _lib.check(L.tg_synth_provider_dev(9, obs.shape[0], 5, _lib.ptr(wave), _lib.ptr(p), _lib.ptr(v), _lib.stream_ptr()))
    return p, v


eng = BatchedMCTS(synth, G, num_simulation=800, tie_seed=1)
a, _, _ = eng.get_action_probs(True); eng.update_with_action(a)          # a warm tree with inherited visits
L.tg_mcts_begin_move(eng._cfg, eng._w(), None, eng._s())
for _ in range(50):
    eng.run_wave()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tb = te = tp = 0.0
n = 100
for _ in range(n):
    ev[0].record()
    L.tg_mcts_wave_begin(eng._cfg, eng._w(), _lib.ptr(eng.obs), eng._s())
    ev[1].record()
    p, v = synth(eng.obs)
    ev[2].record()
    L.tg_mcts_wave_end(eng._cfg, eng._w(), _lib.ptr(p), _lib.ptr(v), None, eng._s())
    ev[3].record()
    torch.cuda.synchronize()
    tb += ev[0].elapsed_time(ev[1]); tp += ev[1].elapsed_time(ev[2]); te += ev[2].elapsed_time(ev[3])
print("wave_begin (select + leaves) %.1f us | provider %.1f us | wave_end (memset + expand) %.1f us" % (tb / n * 1e3, tp / n * 1e3, te / n * 1e3))
hdr = eng.ws.reshape(-1)[:G * 20].reshape(G, 20)
print("replays per game: mean %.3f | terminal %.3f | ties>1 %.3f" % (hdr[:, 17].float().mean().item(), hdr[:, 18].float().mean().item(), hdr[:, 19].float().mean().item()))
