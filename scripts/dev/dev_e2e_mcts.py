"""dev: end-to-end MCTS (1024 games x 800 sims, fused bf16 net, whole wave as one CUDA graph): tcgen05 attention block vs mma path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import make_fused_evaluator
from transgo_b200.mcts import BatchedMCTS
torch.manual_seed(0)
net = TransGoNetwork(board_size=9).cuda().eval()
order = sys.argv[1:] or ["tcgen05", "mma", "tcgen05", "mma"]
for att in order:
    ev = make_fused_evaluator(net, cuda_graph="roots", attention=att)
    eng = BatchedMCTS(ev, 1024, num_simulation=800, tie_seed=1, graph_waves=True)
    for _ in range(3):
        eng.run_wave()
    eng.total_sims()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(); w0 = eng.waves
    act, pi, obs = eng.get_action_probs(True)
    eng.update_with_action(act)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b); n = eng.total_sims()
    print("attention=%-8s %.0f sims/s  %.1f ms  %d waves  %.3f ms/wave" % (att, n / ms * 1e3, ms, eng.waves - w0, ms / (eng.waves - w0)), flush=True)
    del eng, ev
