"""dev: run tg_net_attention_fused once, print per-stage errors against fp32 PyTorch, then time it against GEMM + tg_net_attention."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from transgo_b200 import _lib
from test_net_gpu import _attention_tc_case
L = _lib.lib()


def run(B, P, dbg=True):
    c = _attention_tc_case(B, P, 1000 * B + P)
    out = torch.full_like(c["x"], float("nan"))
    d = torch.full([128 * 192 + 128 * 96 + 128 * 128], float("nan"), device="cuda") if dbg else None
    rc = L.tg_net_attention_fused(_lib.ptr(c["x"]), _lib.ptr(c["w"]), _lib.ptr(c["b"]), c["gamma"], _lib.ptr(c["scale"]), _lib.ptr(c["shift"]),
                                  _lib.ptr(out), B, P, c["dq"], c["C"], _lib.ptr(d), _lib.stream_ptr())
    torch.cuda.synchronize()
    print("B=%d P=%d rc=%d" % (B, P, rc))
    if dbg:
        d1 = d[:128 * 192].reshape(128, 192)[:P]; s = d[128 * 192:128 * 192 + 128 * 96].reshape(128, 96)[:P, :P]; o = d[128 * 192 + 128 * 96:].reshape(128, 128)[:P]
        for name, a, ref in (("projection", d1, c["d1"][0]), ("scores", s, c["s"][0]), ("o", o, c["o"][0])):
            err = (a - ref).abs()
            print("  %-10s max err %.4g  (ref max %.4g, nan %d)  row0[:4] got %s want %s" % (name, float(err.nan_to_num(1e9).max()), float(ref.abs().max()),
                  int(torch.isnan(a).sum()), a[0, :4].tolist(), ref[0, :4].tolist()))
    err = (out.float() - c["ref"]).abs()
    print("  output     max err %.4g (nan %d)" % (float(err.nan_to_num(1e9).max()), int(torch.isnan(out.float()).sum())))
    return c


for B, P in ((1, 81), (5, 81), (3, 96), (4, 49), (700, 81)):
    run(B, P)
# timing at the wave batch
B, P = 4096, 81
c = _attention_tc_case(B, P, 5)
out = torch.empty_like(c["x"])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]


def tc():
    L.tg_net_attention_fused(_lib.ptr(c["x"]), _lib.ptr(c["w"]), _lib.ptr(c["b"]), c["gamma"], _lib.ptr(c["scale"]), _lib.ptr(c["shift"]),
                             _lib.ptr(out), B, P, c["dq"], c["C"], None, _lib.stream_ptr())


def old():
    qkv = torch.nn.functional.linear(c["x"], c["w"], c["b"])
    L.tg_net_attention(_lib.ptr(qkv), _lib.ptr(c["x"]), c["gamma"], _lib.ptr(c["scale"]), _lib.ptr(c["shift"]), _lib.ptr(out), B, P, c["dq"], c["C"], _lib.stream_ptr())


flush = torch.empty([256 << 20], dtype=torch.uint8, device="cuda")
for name, fn in (("tcgen05 fused", tc), ("GEMM + mma.sync", old)):
    for _ in range(3):
        fn()
    t = 0.0
    for _ in range(10):
        flush.zero_()
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
        t += ev[0].elapsed_time(ev[1])
    print("%-16s %.1f us per 4096 positions" % (name, t / 10 * 1e3))
