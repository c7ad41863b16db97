"""dev: fused forward of 4096 positions (CUDA graph replay), tcgen05 attention block vs GEMM + mma.sync."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import make_fused_evaluator
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
net = TransGoNetwork(board_size=9).cuda().eval()
x = (torch.rand(B, 10, 9, 9, device="cuda") > 0.7).float()
res = {}
for att in ("tcgen05", "mma"):
    ev = make_fused_evaluator(net, cuda_graph=True, attention=att)
    for _ in range(3): p, v = ev(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): p, v = ev(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    res[att] = (p.clone(), v.clone())
    print("fused forward, attention=%-8s %.3f ms per %d positions (%.0f TFLOP/s)" % (att, ms, B, 0.457e9 * B / ms / 1e9), flush=True)
print("max |dp| %.2e  max |dv| %.2e between the two attention paths" % (float((res["tcgen05"][0] - res["mma"][0]).abs().max()), float((res["tcgen05"][1] - res["mma"][1]).abs().max())))
