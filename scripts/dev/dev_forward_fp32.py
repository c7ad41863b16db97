"""dev: float32 routes at 4096 positions: the eager module vs the folded schedule (FusedTransGoInference(dtype=float32))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork, make_evaluator
from transgo_b200.fused_net import make_fused_evaluator
torch.manual_seed(0)
B = 4096
net = TransGoNetwork(board_size=9).cuda().eval()
x = (torch.rand(B, 10, 9, 9, device="cuda") > 0.7).float()
for name, ev in (("eager fp32 module (CUDA graph)", make_evaluator(net, cuda_graph=True)),
                 ("folded fp32 schedule (CUDA graph)", make_fused_evaluator(net, dtype=torch.float32, cuda_graph=True)),
                 ("fused bf16 (CUDA graph)", make_fused_evaluator(net, cuda_graph=True))):
    for _ in range(3): out = ev(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): out = ev(x)
    b.record(); torch.cuda.synchronize()
    print("%-36s %.3f ms per %d positions" % (name, a.elapsed_time(b) / 10, B), flush=True)
