"""Net forward timing at the MCTS wave batch (4096 positions) in several inference modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork

def t_ms(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

torch.manual_seed(0)
B = 4096
x = (torch.rand(B, 10, 9, 9, device="cuda") > 0.7).float()
net = TransGoNetwork(board_size=9).cuda().eval()
flops = 0.457e9 * B
def report(name, ms): print("%-46s %8.3f ms  %7.1f TFLOP/s" % (name, ms, flops / ms / 1e9), flush=True)
with torch.no_grad():
    report("fp32 (TF32 conv default)", t_ms(lambda: net.main_prediction(x)))
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    report("fp32 + TF32 matmul", t_ms(lambda: net.main_prediction(x)))
    torch.backends.cudnn.benchmark = True
    report("fp32 + cudnn.benchmark", t_ms(lambda: net.main_prediction(x)))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        report("bf16 autocast", t_ms(lambda: net.main_prediction(x)))
    xcl = x.contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        report("bf16 autocast channels_last input", t_ms(lambda: net.main_prediction(xcl)))
    nb = TransGoNetwork(board_size=9).cuda().eval().to(torch.bfloat16).to(memory_format=torch.channels_last)
    xb = xcl.to(torch.bfloat16)
    report("bf16 weights + channels_last", t_ms(lambda: nb.main_prediction(xb)))
    nh = TransGoNetwork(board_size=9).cuda().eval().half().to(memory_format=torch.channels_last)
    xh = xcl.half()
    report("fp16 weights + channels_last", t_ms(lambda: nh.main_prediction(xh)))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): nb.main_prediction(xb)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        out = nb.main_prediction(xb)
    report("bf16 weights + channels_last + CUDA graph", t_ms(lambda: g.replay()))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        nb.main_prediction(xb); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
    from transgo_b200.fused_net import FusedTransGoInference, make_fused_evaluator
    fn = FusedTransGoInference(net)
    report("fused inference (eager)", t_ms(lambda: fn(x)))
    ev = make_fused_evaluator(net, cuda_graph=True)
    report("fused inference + CUDA graph", t_ms(lambda: ev(x)))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(x); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=80))
