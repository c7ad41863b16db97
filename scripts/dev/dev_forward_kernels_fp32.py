"""dev: eager folded-float32 forwards of 4096 positions (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import FusedTransGoInference
torch.manual_seed(0)
net = TransGoNetwork(board_size=9).cuda().eval()
f = FusedTransGoInference(net, dtype=torch.float32)
x = (torch.rand(4096, 10, 9, 9, device="cuda") > 0.7).float()
for _ in range(2):
    f(x)
torch.cuda.synchronize()
f(x)
torch.cuda.synchronize()
