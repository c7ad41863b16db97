"""Where does the fused forward spend its time at 19x19 (2048 positions = 512 games x 4 leaves)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import FusedTransGoInference
torch.manual_seed(0)
net = TransGoNetwork(board_size=19).cuda().eval()
fn = FusedTransGoInference(net)
x = (torch.rand([2048, 10, 19, 19], device="cuda") > 0.7).float()
for _ in range(3):
    fn(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    fn(x)
b.record(); torch.cuda.synchronize()
print("fused forward 19x19, 2048 positions: %.3f ms" % (a.elapsed_time(b) / 5))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn(x); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
