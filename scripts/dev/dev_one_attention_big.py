import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
L = _lib.lib()
B, P, dq, C = 2048, 361, 32, 128
g = torch.Generator(device="cuda"); g.manual_seed(0)
qkv = (torch.randn([B, P, 2 * dq + C], device="cuda", generator=g) * 0.8).to(torch.bfloat16)
x = torch.randn([B, P, C], device="cuda", generator=g).to(torch.bfloat16)
scale = torch.rand([C], device="cuda", generator=g) + 0.5
shift = torch.randn([C], device="cuda", generator=g) * 0.3
out = torch.empty_like(x)
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(L.tg_net_attention(_lib.ptr(qkv), _lib.ptr(x), 0.6, _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(out), B, P, dq, C, _lib.stream_ptr()))
    b.record(); torch.cuda.synchronize()
    print("attention (361 points, 2048 positions) %.1f us" % (a.elapsed_time(b) * 1e3))
