"""Does splitting the wave batch help the fused forward?  (a) one batch, (b) k sequential chunks (activations fit L2),
(c) k chunks on k streams (tensor-bound convolutions of one chunk overlap the memory-bound glue of another)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200.net import TransGoNetwork
from transgo_b200.fused_net import FusedTransGoInference

torch.manual_seed(0)
net = TransGoNetwork(board_size=9).cuda().eval()
fn = FusedTransGoInference(net)
B = 4096
x = (torch.rand([B, 10, 9, 9], device="cuda") > 0.7).float()


def t_ms(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def graphed(f):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            f()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = f()
    return g.replay


def seq(k):
    xs = x.chunk(k)
    return lambda: [fn(c) for c in xs]


def par(k):
    xs = x.chunk(k)
    streams = [torch.cuda.Stream() for _ in range(k)]

    def f():
        cur = torch.cuda.current_stream()
        outs = []
        for c, s in zip(xs, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(fn(c))
        for s in streams:
            cur.wait_stream(s)
        return outs
    return f


print("one batch of %d, graph: %.3f ms" % (B, t_ms(graphed(lambda: fn(x)))))
for k in (2, 4, 8):
    print("k=%d sequential chunks, graph: %.3f ms" % (k, t_ms(graphed(seq(k)))))
    print("k=%d chunks on %d streams, graph: %.3f ms" % (k, k, t_ms(graphed(par(k)))))
