"""dev: fused forward of 4096 positions, junction tensors walked back to front (L2 tail reuse) vs front to back; alternating."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import make_fused_evaluator
L = _lib.lib()
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
net = TransGoNetwork(board_size=9).cuda().eval()
x = (torch.rand(B, 10, 9, 9, device="cuda") > 0.7).float()
evs = {}
for rev in (0, 1):
    L.tg_net_debug_reverse(rev)
    ev = make_fused_evaluator(net, cuda_graph=True)
    for _ in range(3): ev(x)                       # captured with this setting
    evs[rev] = ev
torch.cuda.synchronize()
tot = {k: [] for k in evs}
for r in range(rounds):
    for rev in ((0, 1) if r % 2 == 0 else (1, 0)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): p, v = evs[rev](x)
        b.record(); torch.cuda.synchronize()
        tot[rev].append(a.elapsed_time(b) / 20)
for rev, t in tot.items():
    print("fused forward, junction order %s: median %.3f ms (min %.3f max %.3f)" % ("back to front" if rev else "front to back", sorted(t)[len(t) // 2], min(t), max(t)), flush=True)
p0 = evs[0](x)[0].clone(); p1 = evs[1](x)[0]
print("identical outputs:", bool(torch.equal(p0, p1)))
