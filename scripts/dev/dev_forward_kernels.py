"""dev: eager fused forwards of 4096 positions (for an ncu launch list: per-kernel time of one forward)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import FusedTransGoInference
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
net = TransGoNetwork(board_size=9).cuda().eval()
f = FusedTransGoInference(net)
x = (torch.rand(B, 10, 9, 9, device="cuda") > 0.7).float()
for _ in range(3):
    f(x)
torch.cuda.synchronize()
print("MARK", flush=True)
f(x)
torch.cuda.synchronize()
