"""dev: tg_net_stem vs staging copy + library convolution (+ junction) at 4096 positions (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
L = _lib.lib()
B = 4096
g = torch.Generator(device="cuda"); g.manual_seed(1)
obs = (torch.rand([B, 10, 9, 9], device="cuda", generator=g) > 0.6).float()
w = (torch.randn([128, 10, 3, 3], device="cuda", generator=g) * 0.2).to(torch.bfloat16)
bias = torch.randn([128], device="cuda", generator=g) * 0.3
s2 = torch.rand([128], device="cuda", generator=g) + 0.5
t2 = torch.randn([128], device="cuda", generator=g) * 0.3
wk = torch.nn.functional.pad(w.permute(0, 2, 3, 1).reshape(128, 9, 10), (0, 6)).contiguous()
out = torch.empty([B, 81, 128], dtype=torch.bfloat16, device="cuda"); out2 = torch.empty_like(out)
x16 = torch.zeros([B, 16, 9, 9], dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
wl = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 6)).contiguous(memory_format=torch.channels_last)
bl = bias.to(torch.bfloat16)
flush = torch.empty([256 << 20], dtype=torch.uint8, device="cuda")


def mine():
    L.tg_net_stem(_lib.ptr(obs), 10, _lib.ptr(wk), _lib.ptr(bias), _lib.ptr(s2), _lib.ptr(t2), _lib.ptr(out), _lib.ptr(out2), B, 9, 128, _lib.stream_ptr())


def mine1():
    L.tg_net_stem(_lib.ptr(obs), 10, _lib.ptr(wk), _lib.ptr(bias), None, None, _lib.ptr(out), None, B, 9, 128, _lib.stream_ptr())


def lib():
    x16[:, :10] = obs
    x = torch.cudnn_convolution_relu(x16, wl, bl, [1, 1], [1, 1], [1, 1], 1)
    a = torch.empty_like(x)
    L.tg_net_add_bn_relu(_lib.ptr(x), None, 1.0, None, _lib.ptr(s2), _lib.ptr(t2), None, _lib.ptr(a), B * 81, 128, _lib.stream_ptr())


for name, fn in (("tg_net_stem (two outputs)", mine), ("tg_net_stem (one output)", mine1), ("copy + cudnn_convolution_relu + junction", lib)):
    for _ in range(3): fn()
    t = 0.0
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        t += a.elapsed_time(b)
    print("%-44s %.1f us" % (name, t / 10 * 1e3))
