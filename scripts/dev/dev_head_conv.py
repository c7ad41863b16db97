"""dev: tg_net_conv3x3_head vs the library convolution at 4096 positions (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
L = _lib.lib()
B, nout = 4096, 4
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn([B, 81, 128], device="cuda", generator=g).to(torch.bfloat16)
w = (torch.randn([nout, 128, 3, 3], device="cuda", generator=g) * 0.05).to(torch.bfloat16)
w8 = torch.zeros([8, 9, 128], dtype=torch.bfloat16, device="cuda"); w8[:nout] = w.permute(0, 2, 3, 1).reshape(nout, 9, 128)
b8 = torch.zeros([8], dtype=torch.float32, device="cuda")
out = torch.empty([B, 81, nout], dtype=torch.bfloat16, device="cuda")
xc = x.reshape(B, 9, 9, 128).permute(0, 3, 1, 2)          # channels_last view
wl = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, 8 - nout)).contiguous(memory_format=torch.channels_last)
bl = torch.zeros([8], dtype=torch.bfloat16, device="cuda")
flush = torch.empty([256 << 20], dtype=torch.uint8, device="cuda")


def mine():
    L.tg_net_conv3x3_head(_lib.ptr(x), _lib.ptr(w8), _lib.ptr(b8), _lib.ptr(out), B, 9, 128, nout, _lib.stream_ptr())


def lib():
    torch.cudnn_convolution_relu(xc, wl, bl, [1, 1], [1, 1], [1, 1], 1)


for name, fn in (("tg_net_conv3x3_head", mine), ("cudnn_convolution_relu (8 channels)", lib)):
    for _ in range(3): fn()
    t = 0.0
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        t += a.elapsed_time(b)
    print("%-40s %.1f us" % (name, t / 10 * 1e3))
