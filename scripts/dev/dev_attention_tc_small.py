"""dev: a few small launches of tg_net_attention_fused (for compute-sanitizer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from transgo_b200 import _lib
from test_net_gpu import _attention_tc_case
L = _lib.lib()
for B, P in ((3, 81), (600, 81), (2, 96), (2, 7)):
    c = _attention_tc_case(B, P, 1000 * B + P)
    out = torch.zeros_like(c["x"])
    rc = L.tg_net_attention_fused(_lib.ptr(c["x"]), _lib.ptr(c["w"]), _lib.ptr(c["b"]), c["gamma"], _lib.ptr(c["scale"]), _lib.ptr(c["shift"]),
                                  _lib.ptr(out), B, P, c["dq"], c["C"], None, _lib.stream_ptr())
    torch.cuda.synchronize()
    err = (out.float() - c["ref"]).abs().max().item()
    print("B=%d P=%d rc=%d max err %.4f" % (B, P, rc, err), flush=True)
