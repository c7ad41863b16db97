"""dev: fused forward of 4096 positions (CUDA graph replay), tcgen05 attention block vs GEMM + mma.sync, alternating so that
clock / power drift hits both alike."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import make_fused_evaluator
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
net = TransGoNetwork(board_size=9).cuda().eval()
with torch.no_grad():
    for k, p in net.named_parameters():
        if k.endswith("gamma"):
            p.fill_(0.4)
x = (torch.rand(B, 10, 9, 9, device="cuda") > 0.7).float()
evs = {att: make_fused_evaluator(net, cuda_graph=True, attention=att) for att in ("tcgen05", "mma")}
for ev in evs.values():
    for _ in range(3): ev(x)
torch.cuda.synchronize()
tot = {k: [] for k in evs}
for r in range(rounds):
    for att in (("tcgen05", "mma") if r % 2 == 0 else ("mma", "tcgen05")):
        ev = evs[att]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): p, v = ev(x)
        b.record(); torch.cuda.synchronize()
        tot[att].append(a.elapsed_time(b) / 20)
for att, t in tot.items():
    print("fused forward, attention=%-8s median %.3f ms (min %.3f max %.3f) per %d positions" % (att, sorted(t)[len(t) // 2], min(t), max(t), B), flush=True)
p1, v1 = evs["tcgen05"](x); p1 = p1.clone(); v1 = v1.clone()
p2, v2 = evs["mma"](x)
print("max |dp| %.2e  max |dv| %.2e between the two attention paths" % (float((p1 - p2).abs().max()), float((v1 - v2).abs().max())))
