"""dev: eager fused forwards of 2048 positions of 19x19 (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import FusedTransGoInference
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
net = TransGoNetwork(board_size=19).cuda().eval()
f = FusedTransGoInference(net)
x = (torch.rand(B, 10, 19, 19, device="cuda") > 0.7).float()
for _ in range(2):
    f(x)
torch.cuda.synchronize()
f(x)
torch.cuda.synchronize()
