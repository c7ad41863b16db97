"""GPU numerics tests of the inference path of the policy/value net: the hand-written junction kernel against a plain
PyTorch fp32 restatement of the same op (bit-equal: same rounding points), and the fused bf16 forward against the fp32
module (tolerance stated below; bf16 storage has 8 mantissa bits)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _perturbed_model(seed=0, features=128):
    from transgo_b200.net import TransGoNetwork
    torch.manual_seed(seed)
    m = TransGoNetwork(board_size=9, num_features=features).cuda().eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.7, 1.3); mod.bias.normal_(0, 0.1)
        for k, p in m.named_parameters():
            if k.endswith("gamma"):
                p.fill_(0.4)
    return m


@pytest.mark.parametrize("rows,C,with_y,want_sum,want_act,with_bias", [(81 * 7, 128, True, True, True, True), (81 * 3, 128, False, False, True, False),
                                                                        (1000, 64, True, True, False, False), (5, 8, True, False, True, True)])
def test_junction_kernel_matches_fp32_reference(rows, C, with_y, want_sum, want_act, with_bias):
    from transgo_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda"); g.manual_seed(rows + C)
    x = torch.randn([rows, C], device="cuda", generator=g).to(torch.bfloat16)
    y = torch.randn([rows, C], device="cuda", generator=g).to(torch.bfloat16) if with_y else None
    scale = torch.rand([C], device="cuda", generator=g) + 0.5
    shift = torch.randn([C], device="cuda", generator=g)
    alpha = 0.37
    ybias = torch.randn([C], device="cuda", generator=g) if with_bias else None
    s_out = torch.empty_like(x) if want_sum else None
    a_out = torch.empty_like(x) if want_act else None
    _lib.check(L.tg_net_add_bn_relu(_lib.ptr(x), _lib.ptr(y), alpha, _lib.ptr(ybias), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(s_out),
                                    _lib.ptr(a_out), rows, C, _lib.stream_ptr()))
    yb = (y.float() + ybias) if (with_y and with_bias) else (y.float() if with_y else None)
    ref_sum = x.float() + (torch.tensor(alpha, device="cuda") * yb if with_y else 0.0)
    ref_sum = ref_sum.to(torch.bfloat16)
    ref_act = torch.relu(ref_sum.float() * scale + shift).to(torch.bfloat16)
    if want_sum:
        assert torch.equal(s_out, ref_sum)
    if want_act:
        assert torch.equal(a_out, ref_act)


@pytest.mark.parametrize("B,P", [(5, 81), (300, 81), (3, 96), (4, 49), (2, 1), (3, 361), (40, 361), (2, 97), (2, 368), (3, 200)])
def test_attention_kernel_matches_fp32_reference(B, P):
    """tg_net_attention against a plain PyTorch fp32 restatement of model.py:299-314 on the same bf16 inputs.
    Tolerance: |d| <= 2e-2 + 2e-2 * |ref| (probabilities and outputs are rounded to bf16: 8 mantissa bits)."""
    from transgo_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda"); g.manual_seed(B * 100 + P)
    dq, C = 32, 128
    qkv = (torch.randn([B, P, 2 * dq + C], device="cuda", generator=g) * 0.8).to(torch.bfloat16)
    x = torch.randn([B, P, C], device="cuda", generator=g).to(torch.bfloat16)
    scale = torch.rand([C], device="cuda", generator=g) + 0.5
    shift = torch.randn([C], device="cuda", generator=g) * 0.3
    gamma = 0.6
    out = torch.full_like(x, float("nan"))
    _lib.check(L.tg_net_attention(_lib.ptr(qkv), _lib.ptr(x), gamma, _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(out),
                                  B, P, dq, C, _lib.stream_ptr()))
    q, k, v = qkv[..., :dq].float(), qkv[..., dq:2 * dq].float(), qkv[..., 2 * dq:].float()
    att = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)                 # [B, P(query i), P(key j)]
    o = torch.bmm(att.transpose(1, 2), v)                                          # o[j] = sum_i att[i, j] v[i]
    ref = torch.relu((x.float() + gamma * o) * scale + shift)
    assert not torch.isnan(out.float()).any()
    err = (out.float() - ref).abs()
    assert bool((err <= 2e-2 + 2e-2 * ref.abs()).all()), float(err.max())
    assert L.tg_net_attention(_lib.ptr(qkv), _lib.ptr(x), gamma, _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(out),
                              B, 369, dq, C, _lib.stream_ptr()) != 0         # unsupported size is refused, not mangled


def _attention_tc_case(B, P, seed):
    """inputs and fp32 stage-by-stage expectations of tg_net_attention_fused on the same bf16 operands"""
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    dq, C = 32, 128
    x = torch.randn([B, P, C], device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn([2 * dq + C, C], device="cuda", generator=g) * 0.09).to(torch.bfloat16)
    b = (torch.randn([2 * dq + C], device="cuda", generator=g) * 0.2).to(torch.bfloat16)
    scale = torch.rand([C], device="cuda", generator=g) + 0.5
    shift = torch.randn([C], device="cuda", generator=g) * 0.3
    gamma = 0.6
    d1 = x.float() @ w.float().t() + b.float()                                    # projection + bias (fp32 accumulate)
    qkv = d1.to(torch.bfloat16).float()                                           # parked as bf16
    q, k, v = qkv[..., :dq], qkv[..., dq:2 * dq], qkv[..., 2 * dq:]
    s = torch.bmm(q, k.transpose(1, 2))                                           # [B, i, j]
    att = torch.softmax(s, dim=-1).to(torch.bfloat16).float()
    o = torch.bmm(att.transpose(1, 2), v)                                         # o[j] = sum_i att[i, j] v[i]
    ref = torch.relu((x.float() + gamma * o.to(torch.bfloat16).float()) * scale + shift)
    return dict(x=x, w=w, b=b, scale=scale, shift=shift, gamma=gamma, d1=d1, s=s, o=o, ref=ref, dq=dq, C=C)


@pytest.mark.parametrize("B,P", [(1, 81), (5, 81), (700, 81), (3, 96), (4, 49), (2, 1)])
def test_attention_tcgen05_kernel_stage_by_stage(B, P):
    """tg_net_attention_fused (tcgen05: projection, scores, att^T . v in tensor memory) against fp32 PyTorch on the same bf16
    operands.  The accumulator tiles of position 0 are dumped and checked one contraction at a time (a wrong shared-memory
    descriptor shows up in the first stage that uses it).  Tolerances: fp32 accumulators |d| <= 2e-3 + 2e-3 |ref| (operands are
    exact bf16 products, only the summation order differs; the scores / o stages also see bf16-rounded q, k, v / p);
    output |d| <= 2e-2 + 2e-2 |ref| (bf16 rounding of p, o and the result: 8 mantissa bits)."""
    from transgo_b200 import _lib
    L = _lib.lib()
    c = _attention_tc_case(B, P, 1000 * B + P)
    out = torch.full_like(c["x"], float("nan"))
    dbg = torch.full([128 * 192 + 128 * 96 + 128 * 128], float("nan"), device="cuda")
    _lib.check(L.tg_net_attention_fused(_lib.ptr(c["x"]), _lib.ptr(c["w"]), _lib.ptr(c["b"]), c["gamma"], _lib.ptr(c["scale"]),
                                        _lib.ptr(c["shift"]), _lib.ptr(out), B, P, c["dq"], c["C"], _lib.ptr(dbg), _lib.stream_ptr()),
               "tg_net_attention_fused")
    torch.cuda.synchronize()
    d1 = dbg[:128 * 192].reshape(128, 192)[:P]
    s = dbg[128 * 192:128 * 192 + 128 * 96].reshape(128, 96)[:P, :P]
    o = dbg[128 * 192 + 128 * 96:].reshape(128, 128)[:P]

    def close(a, ref, tol, what):
        err = (a - ref).abs()
        assert bool((err <= tol + tol * ref.abs()).all()), (what, float(err.max()), float(ref.abs().max()))
    close(d1, c["d1"][0], 2e-3, "projection")
    close(s, c["s"][0], 3e-2, "scores")                   # q, k rounded to bf16 on both sides; fp32 sums of 32 products
    close(o, c["o"][0], 2e-2, "att^T . v")                # p rounded to bf16 after an approximate exp2
    assert not torch.isnan(out.float()).any()
    close(out.float(), c["ref"], 2e-2, "output")
    assert L.tg_net_attention_fused(_lib.ptr(c["x"]), _lib.ptr(c["w"]), _lib.ptr(c["b"]), c["gamma"], _lib.ptr(c["scale"]),
                                    _lib.ptr(c["shift"]), _lib.ptr(out), B, 97, c["dq"], c["C"], None, _lib.stream_ptr()) != 0


def test_attention_tcgen05_second_output_is_the_next_blocks_activation():
    """tg_net_attention_fused2: out2 = relu(out * scale2 + shift2) computed from the bf16-rounded out -- bit-equal to running the
    junction kernel (tg_net_add_bn_relu) on out, and out itself unchanged by asking for it."""
    from transgo_b200 import _lib
    L = _lib.lib()
    B, P = 300, 81
    c = _attention_tc_case(B, P, 4242)
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    s2 = torch.rand([128], device="cuda", generator=g) + 0.5
    t2 = torch.randn([128], device="cuda", generator=g) * 0.3
    out = torch.empty_like(c["x"]); out2 = torch.full_like(c["x"], float("nan")); out_plain = torch.empty_like(c["x"])
    _lib.check(L.tg_net_attention_fused2(_lib.ptr(c["x"]), _lib.ptr(c["w"]), _lib.ptr(c["b"]), c["gamma"], _lib.ptr(c["scale"]), _lib.ptr(c["shift"]),
                                         _lib.ptr(out), _lib.ptr(s2), _lib.ptr(t2), _lib.ptr(out2), B, P, c["dq"], c["C"], None, _lib.stream_ptr()))
    _lib.check(L.tg_net_attention_fused(_lib.ptr(c["x"]), _lib.ptr(c["w"]), _lib.ptr(c["b"]), c["gamma"], _lib.ptr(c["scale"]), _lib.ptr(c["shift"]),
                                        _lib.ptr(out_plain), B, P, c["dq"], c["C"], None, _lib.stream_ptr()))
    ref2 = torch.empty_like(out)
    _lib.check(L.tg_net_add_bn_relu(_lib.ptr(out), None, 1.0, None, _lib.ptr(s2), _lib.ptr(t2), None, _lib.ptr(ref2), B * P, 128, _lib.stream_ptr()))
    assert torch.equal(out, out_plain) and torch.equal(out2, ref2)


@pytest.mark.parametrize("B,P", [(40, 81), (6, 361)])
def test_attention_mma_second_output_is_the_next_blocks_activation(B, P):
    """tg_net_attention2 (mma.sync kernels, both board sizes): out2 bit-equal to the junction kernel applied to out."""
    from transgo_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda"); g.manual_seed(B + P)
    qkv = (torch.randn([B, P, 192], device="cuda", generator=g) * 0.8).to(torch.bfloat16)
    x = torch.randn([B, P, 128], device="cuda", generator=g).to(torch.bfloat16)
    sc, sh, s2, t2 = (torch.rand([128], device="cuda", generator=g) + 0.5 for _ in range(4))
    out = torch.empty_like(x); out2 = torch.full_like(x, float("nan")); plain = torch.empty_like(x); ref2 = torch.empty_like(x)
    _lib.check(L.tg_net_attention2(_lib.ptr(qkv), _lib.ptr(x), 0.6, _lib.ptr(sc), _lib.ptr(sh), _lib.ptr(out), _lib.ptr(s2), _lib.ptr(t2), _lib.ptr(out2),
                                   B, P, 32, 128, _lib.stream_ptr()))
    _lib.check(L.tg_net_attention(_lib.ptr(qkv), _lib.ptr(x), 0.6, _lib.ptr(sc), _lib.ptr(sh), _lib.ptr(plain), B, P, 32, 128, _lib.stream_ptr()))
    _lib.check(L.tg_net_add_bn_relu(_lib.ptr(out), None, 1.0, None, _lib.ptr(s2), _lib.ptr(t2), None, _lib.ptr(ref2), B * P, 128, _lib.stream_ptr()))
    assert torch.equal(out, plain) and torch.equal(out2, ref2)


def test_attention_tcgen05_kernel_agrees_with_the_mma_sync_kernel():
    """the tcgen05 block against the round-1 pair (library GEMM + tg_net_attention) on the same inputs: both round q, k, v, p
    and o to bf16 at the same points, so they differ by summation order only (|d| <= 2e-2 + 2e-2 |ref|, in practice bf16 ulps)."""
    from transgo_b200 import _lib
    L = _lib.lib()
    B, P = 64, 81
    c = _attention_tc_case(B, P, 77)
    out_tc = torch.empty_like(c["x"]); out_mma = torch.empty_like(c["x"])
    _lib.check(L.tg_net_attention_fused(_lib.ptr(c["x"]), _lib.ptr(c["w"]), _lib.ptr(c["b"]), c["gamma"], _lib.ptr(c["scale"]),
                                        _lib.ptr(c["shift"]), _lib.ptr(out_tc), B, P, c["dq"], c["C"], None, _lib.stream_ptr()))
    qkv = torch.nn.functional.linear(c["x"], c["w"], c["b"])
    _lib.check(L.tg_net_attention(_lib.ptr(qkv), _lib.ptr(c["x"]), c["gamma"], _lib.ptr(c["scale"]), _lib.ptr(c["shift"]), _lib.ptr(out_mma),
                                  B, P, c["dq"], c["C"], _lib.stream_ptr()))
    err = (out_tc.float() - out_mma.float()).abs()
    assert bool((err <= 2e-2 + 2e-2 * out_mma.float().abs()).all()), float(err.max())


@pytest.mark.parametrize("B,with_second", [(1, True), (9, False), (700, True)])
def test_stem_kernel_matches_fp32_reference(B, with_second):
    """tg_net_stem (planes float32 [B,10,81] -> relu(conv3x3 + bias) NHWC bf16, optional second output relu(out * s2 + t2)) against
    conv2d in fp32 on the same bf16-rounded operands.  Tolerance |d| <= 1e-2 + 1e-2 |ref|; the second output is bit-equal to
    the junction kernel applied to the first."""
    from transgo_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda"); g.manual_seed(31 * B)
    obs = (torch.rand([B, 10, 9, 9], device="cuda", generator=g) > 0.6).float()
    obs[:, 3] = torch.rand([B, 9, 9], device="cuda", generator=g)                  # not only 0 / 1 inputs
    w = (torch.randn([128, 10, 3, 3], device="cuda", generator=g) * 0.2).to(torch.bfloat16)
    bias = torch.randn([128], device="cuda", generator=g) * 0.3
    s2 = torch.rand([128], device="cuda", generator=g) + 0.5
    t2 = torch.randn([128], device="cuda", generator=g) * 0.3
    wk = torch.nn.functional.pad(w.permute(0, 2, 3, 1).reshape(128, 9, 10), (0, 6)).contiguous()
    out = torch.full([B, 81, 128], float("nan"), dtype=torch.bfloat16, device="cuda")
    out2 = torch.full_like(out, float("nan")) if with_second else None
    _lib.check(L.tg_net_stem(_lib.ptr(obs), 10, _lib.ptr(wk), _lib.ptr(bias), _lib.ptr(s2) if with_second else None, _lib.ptr(t2) if with_second else None,
                             _lib.ptr(out), _lib.ptr(out2), B, 9, 128, _lib.stream_ptr()), "tg_net_stem")
    ref = torch.relu(torch.nn.functional.conv2d(obs.to(torch.bfloat16).float(), w.float(), bias, padding=1)).permute(0, 2, 3, 1).reshape(B, 81, 128)
    assert not torch.isnan(out.float()).any()
    err = (out.float() - ref).abs()
    assert bool((err <= 1e-2 + 1e-2 * ref.abs()).all()), float(err.max())
    if with_second:
        ref2 = torch.empty_like(out)
        _lib.check(L.tg_net_add_bn_relu(_lib.ptr(out), None, 1.0, None, _lib.ptr(s2), _lib.ptr(t2), None, _lib.ptr(ref2), B * 81, 128, _lib.stream_ptr()))
        assert torch.equal(out2, ref2)


@pytest.mark.parametrize("B,nout", [(1, 2), (7, 4), (700, 4), (300, 2), (3, 3), (2, 1)])
def test_head_conv_kernel_matches_fp32_reference(B, nout):
    """tg_net_conv3x3_head (3x3, 128 -> nout channels, bias, ReLU; zero-bordered frames in shared memory + mma.sync) against
    torch.nn.functional.conv2d in fp32 on the same bf16 operands.  Tolerance |d| <= 1e-2 + 1e-2 |ref| (fp32 accumulation of 1152
    bf16 products on both sides; the result is rounded to bf16)."""
    from transgo_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda"); g.manual_seed(10 * B + nout)
    x = torch.randn([B, 81, 128], device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn([nout, 128, 3, 3], device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn([nout], device="cuda", generator=g) * 0.3
    w8 = torch.zeros([8, 9, 128], dtype=torch.bfloat16, device="cuda"); w8[:nout] = w.permute(0, 2, 3, 1).reshape(nout, 9, 128)
    b8 = torch.zeros([8], dtype=torch.float32, device="cuda"); b8[:nout] = bias
    out = torch.full([B, 81, nout], float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(L.tg_net_conv3x3_head(_lib.ptr(x), _lib.ptr(w8), _lib.ptr(b8), _lib.ptr(out), B, 9, 128, nout, _lib.stream_ptr()), "tg_net_conv3x3_head")
    xi = x.float().reshape(B, 9, 9, 128).permute(0, 3, 1, 2)
    ref = torch.relu(torch.nn.functional.conv2d(xi, w.float(), bias, padding=1)).permute(0, 2, 3, 1).reshape(B, 81, nout)
    assert not torch.isnan(out.float()).any()
    err = (out.float() - ref).abs()
    assert bool((err <= 1e-2 + 1e-2 * ref.abs()).all()), float(err.max())
    assert L.tg_net_conv3x3_head(_lib.ptr(x), _lib.ptr(w8), _lib.ptr(b8), _lib.ptr(out), B, 19, 128, nout, _lib.stream_ptr()) != 0
    assert L.tg_net_conv3x3_head(_lib.ptr(x), _lib.ptr(w8), _lib.ptr(b8), _lib.ptr(out), B, 9, 128, 5, _lib.stream_ptr()) != 0


@pytest.mark.parametrize("attention,head_conv,stem_conv", [("tcgen05", "mma", "mma"), ("mma", "library", "library"), ("tcgen05", "library", "mma"),
                                                           ("mma", "mma", "library")])
def test_fused_forward_matches_fp32_module(attention, head_conv, stem_conv):
    """Tolerance: policy max |dp| <= 4e-3 (probabilities), value / ownership max |dv| <= 4e-2 -- bf16 activations through
    26 convolution layers; argmax of the policy must agree wherever the fp32 margin exceeds the tolerance."""
    from transgo_b200.fused_net import FusedTransGoInference
    m = _perturbed_model()
    fused = FusedTransGoInference(m, attention=attention, head_conv=head_conv, stem_conv=stem_conv)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    obs = (torch.rand([257, 10, 9, 9], device="cuda", generator=g) > 0.75).float()
    with torch.no_grad():
        p0, v0, o0 = m.main_prediction(obs)
    p1, v1, o1 = fused(obs)
    assert p1.shape == p0.shape and v1.shape == (257,) and o1.shape == o0.shape
    assert float((p1.sum(1) - 1).abs().max()) < 1e-4
    assert float((p1 - p0).abs().max()) <= 4e-3, float((p1 - p0).abs().max())
    assert float((v1 - v0.reshape(-1)).abs().max()) <= 4e-2 and float((o1 - o0).abs().max()) <= 4e-2
    top2 = p0.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 8e-3
    assert bool((p1.argmax(1)[clear] == p0.argmax(1)[clear]).all())


def test_folded_forward_in_float32_matches_the_module_closely():
    """The same schedule with float32 storage (BatchNorm folded, NHWC, library kernels for the pieces the bf16 kernels cover): the
    same-precision-as-the-reference route.  Tolerance: policy 2e-4, value / ownership 2e-3 (folding reorders float32 roundings; the
    convolutions run on TF32 tensor cores in both, as PyTorch does by default)."""
    from transgo_b200.fused_net import FusedTransGoInference
    m = _perturbed_model(4)
    fused = FusedTransGoInference(m, dtype=torch.float32)
    g = torch.Generator(device="cuda"); g.manual_seed(6)
    obs = (torch.rand([130, 10, 9, 9], device="cuda", generator=g) > 0.75).float()
    with torch.no_grad():
        p0, v0, o0 = m.main_prediction(obs)
    p1, v1, o1 = fused(obs)
    assert p1.dtype == torch.float32
    assert float((p1 - p0).abs().max()) <= 2e-4, float((p1 - p0).abs().max())
    assert float((v1 - v0.reshape(-1)).abs().max()) <= 2e-3 and float((o1 - o0).abs().max()) <= 2e-3


def test_search_runs_on_the_fused_evaluator_and_graph_replay_is_stable():
    from transgo_b200.fused_net import make_fused_evaluator
    from transgo_b200.mcts import BatchedMCTS
    m = _perturbed_model(1)
    ev = make_fused_evaluator(m, cuda_graph=True)
    eng = BatchedMCTS(ev, 64, num_simulation=32, tie_seed=2)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    noise = eng.make_noise(generator=g)
    a1, pi1, _ = eng.get_action_probs(True, noise=noise, u=torch.full([64], 0.5, dtype=torch.float64, device="cuda"))
    c1 = eng.counts.clone()
    ev2 = make_fused_evaluator(m, cuda_graph=False)
    eng2 = BatchedMCTS(ev2, 64, num_simulation=32, tie_seed=2)
    a2, pi2, _ = eng2.get_action_probs(True, noise=noise, u=torch.full([64], 0.5, dtype=torch.float64, device="cuda"))
    assert torch.equal(c1, eng2.counts) and torch.equal(a1, a2)          # graph replay == eager, bit for bit
    assert bool((eng.root_n >= 32).all())
    eng.update_with_action(a1)


def test_fused_forward_on_19x19():
    """19x19 runs the two-pass attention kernel (361 points).  Same tolerances as the 9x9 test."""
    from transgo_b200.fused_net import FusedTransGoInference
    from transgo_b200.net import TransGoNetwork
    torch.manual_seed(5)
    m = TransGoNetwork(board_size=19).cuda().eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5)
        for k, p in m.named_parameters():
            if k.endswith("gamma"):
                p.fill_(0.3)
    fused = FusedTransGoInference(m)
    g = torch.Generator(device="cuda"); g.manual_seed(8)
    obs = (torch.rand([5, 10, 19, 19], device="cuda", generator=g) > 0.75).float()
    with torch.no_grad():
        p0, v0, o0 = m.main_prediction(obs)
    p1, v1, o1 = fused(obs)
    assert p1.shape == (5, 362) and o1.shape == (5, 361)
    assert float((p1 - p0).abs().max()) <= 4e-3 and float((v1 - v0.reshape(-1)).abs().max()) <= 4e-2 and float((o1 - o0).abs().max()) <= 4e-2


def test_refresh_updates_weights_inside_captured_graphs():
    """set_weights + refresh() must reach forwards that were already captured as CUDA graphs (folded tensors are updated in
    place; gamma is folded into the value projection, so no host scalar is baked into a graph)."""
    from transgo_b200.fused_net import FusedTransGoInference, make_fused_evaluator
    m = _perturbed_model(2)
    ev = make_fused_evaluator(m, cuda_graph=True)
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    obs = (torch.rand([64, 10, 9, 9], device="cuda", generator=g) > 0.75).float()
    p_old = ev(obs)[0].clone()
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith("gamma"):
                p.fill_(-0.7)
            elif "fc_act" in k or "conv_act" in k:
                p.mul_(1.5)
    ev.refresh()
    p_new = ev(obs)[0].clone()                                   # replayed graph, new weights
    p_ref = FusedTransGoInference(m)(obs)[0]                     # folded from scratch
    assert torch.equal(p_new, p_ref) and not torch.equal(p_new, p_old)
