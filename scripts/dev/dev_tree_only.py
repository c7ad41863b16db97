"""dev: tree-only MCTS throughput at the C3 size (1024 games x 800 sims, synthetic provider, CUDA-graphed waves), speculative
child prefetch on / off, alternating."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
from transgo_b200.mcts import BatchedMCTS, _declare
L = _lib.lib(); _declare(L)
G = 1024
pol = torch.empty([G * 4, 82], dtype=torch.float32, device="cuda"); val = torch.empty([G * 4], dtype=torch.float32, device="cuda")
pol_r = torch.empty([G, 82], dtype=torch.float32, device="cuda"); val_r = torch.empty([G], dtype=torch.float32, device="cuda")


def run(prefetch, moves=3):
    L.tg_mcts_debug_prefetch(prefetch)
    wave = torch.zeros([2], dtype=torch.int64, device="cuda")

    def synth(obs):
        p, v = (pol, val) if obs.shape[0] == G * 4 else (pol_r, val_r)
        _lib.check(L.tg_synth_provider_dev(9, obs.shape[0], 5, _lib.ptr(wave), _lib.ptr(p), _lib.ptr(v), _lib.stream_ptr()))
        return p, v
    eng = BatchedMCTS(synth, G, num_simulation=800, tie_seed=1, graph_waves=True)
    for _ in range(3):
        eng.run_wave()
    eng.total_sims()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); w0 = eng.waves
    for _ in range(moves):
        act, pi, obs = eng.get_action_probs(True)
        eng.update_with_action(act)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b); n = eng.total_sims()
    cs = int(eng.counts.sum().item())
    return n / ms * 1e3, ms / (eng.waves - w0) * 1e3, cs


for rep in range(3):
    for pf in ((1, 0) if rep % 2 == 0 else (0, 1)):
        r, us, cs = run(pf)
        print("prefetch=%d  %.3e sims/s  %.1f us per wave  (visit-count checksum %d)" % (pf, r, us, cs), flush=True)
