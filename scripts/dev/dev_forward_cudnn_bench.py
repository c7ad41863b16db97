"""dev: does cudnn.benchmark pick faster convolution kernels for the fused forward?  (separate processes would be cleaner; here the
flag is flipped before each evaluator is built and warmed up, and the two are timed alternately)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import make_fused_evaluator
torch.manual_seed(0)
B = 4096
net = TransGoNetwork(board_size=9).cuda().eval()
x = (torch.rand(B, 10, 9, 9, device="cuda") > 0.7).float()
evs = {}
for flag in (False, True):
    torch.backends.cudnn.benchmark = flag
    ev = make_fused_evaluator(net, cuda_graph=True)
    for _ in range(3): ev(x)
    evs[flag] = ev
torch.cuda.synchronize()
tot = {k: [] for k in evs}
for r in range(8):
    for flag in ((False, True) if r % 2 == 0 else (True, False)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): evs[flag](x)
        b.record(); torch.cuda.synchronize()
        tot[flag].append(a.elapsed_time(b) / 20)
for flag, t in tot.items():
    print("cudnn.benchmark=%s: median %.3f ms (min %.3f max %.3f)" % (flag, sorted(t)[len(t) // 2], min(t), max(t)), flush=True)
