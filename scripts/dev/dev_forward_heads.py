"""dev: fused forward of 4096 positions with the hand-written head convolutions vs the library ones; alternating."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import make_fused_evaluator
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
net = TransGoNetwork(board_size=9).cuda().eval()
x = (torch.rand(B, 10, 9, 9, device="cuda") > 0.7).float()
evs = {hc: make_fused_evaluator(net, cuda_graph=True, head_conv=hc, stem_conv=hc) for hc in ("mma", "library")}
for ev in evs.values():
    for _ in range(3): ev(x)
torch.cuda.synchronize()
tot = {k: [] for k in evs}
for r in range(rounds):
    for hc in (("mma", "library") if r % 2 == 0 else ("library", "mma")):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): p, v = evs[hc](x)
        b.record(); torch.cuda.synchronize()
        tot[hc].append(a.elapsed_time(b) / 20)
for hc, t in tot.items():
    print("fused forward, head + stem convolutions=%-8s median %.3f ms (min %.3f max %.3f)" % (hc, sorted(t)[len(t) // 2], min(t), max(t)), flush=True)
p1, v1 = evs["mma"](x); p1 = p1.clone(); v1 = v1.clone()
p2, v2 = evs["library"](x)
print("max |dp| %.2e  max |dv| %.2e between the two" % (float((p1 - p2).abs().max()), float((v1 - v2).abs().max())))
