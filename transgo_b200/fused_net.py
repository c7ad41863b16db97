"""Inference-only forward of the TransGo net (model.py:41-114) for the search's wave batches.

Same weights, same maths, re-scheduled for a B200: the tensor-core work stays in PyTorch library calls (cuDNN 3x3 convolutions
with bias+ReLU epilogue, cuBLAS GEMMs for the 1x1 convolutions / attention / heads), eval-mode BatchNorm is folded
 - into the PRECEDING convolution where one exists (stem, the second BN of every residual block, both head convs), and
 - into a per-channel scale/shift otherwise, applied by ONE hand-written kernel per residual junction
   (tg_net_add_bn_relu: residual add / gamma*attention + x, next BatchNorm, ReLU in a single NHWC bf16 pass);
 - the four self-attention blocks (scores, softmax, weighted sum, junction) run as ONE hand-written kernel per block
   -- up to 96 points (9x9) INCLUDING the q|k|v projection: tg_net_attention_fused, a tcgen05 kernel (UMMA from shared-memory
   operand panels, accumulators in tensor memory; csrc/tg_net_tc.cu); up to 368 points (19x19) tg_net_attention (mma.sync bf16,
   two passes over the scores) fed by a single q|k|v GEMM; larger boards use library batched GEMMs.
Activations are NHWC bf16 end to end, so the attention blocks read them as [B, P, C] views without layout copies.

``FusedTransGoInference(model)`` snapshots a ``TransGoNetwork``; ``refresh()`` re-folds after ``set_weights``.
Numerics: bf16 storage with fp32 accumulation (library kernels) -- compared against the fp32 module in tests/test_net_gpu.py.
"""
import torch
import torch.nn.functional as F

from . import _lib
from .net import _PreActResidual, _SpatialAttention


def _fold(bn):
    """eval-mode BatchNorm2d -> (scale, shift) float32 with y = x * scale + shift"""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def _conv_bn(conv, bn, dtype, pad_in=None):
    """conv followed by BatchNorm -> (weight [O,I,3,3] channels_last in `dtype`, bias) with the BN folded in"""
    w, b = conv.weight.detach().float(), conv.bias.detach().float()
    if bn is not None:
        s, t = _fold(bn)
        w = w * s[:, None, None, None]
        b = b * s + t
    if pad_in is not None and w.shape[1] < pad_in:
        w = F.pad(w, (0, 0, 0, 0, 0, pad_in - w.shape[1]))
    return w.to(dtype).contiguous(memory_format=torch.channels_last), b.to(dtype).contiguous()


def _conv_head(conv, bn, dtype):
    """a head convolution (3x3, few output channels) + BatchNorm for tg_net_conv3x3_head: (weight [8, 9, I] = [output channel,
    tap ky * 3 + kx, input channel] in `dtype`, rows beyond the real output channels zero; bias float32 [8]; channel count)"""
    w, b = conv.weight.detach().float(), conv.bias.detach().float()
    s, t = _fold(bn)
    w = w * s[:, None, None, None]
    b = b * s + t
    O, I = w.shape[0], w.shape[1]
    w8 = torch.zeros([8, 9, I], dtype=torch.float32, device=w.device)
    w8[:O] = w.permute(0, 2, 3, 1).reshape(O, 9, I)
    b8 = torch.zeros([8], dtype=torch.float32, device=w.device)
    b8[:O] = b
    return w8.to(dtype).contiguous(), b8.contiguous()


def _copy_tree(old, new):
    """copy every tensor of the nested structure `new` into the same place of `old` (same layout by construction)"""
    if isinstance(old, torch.Tensor):
        old.copy_(new)
    elif isinstance(old, dict):
        for k in old:
            _copy_tree(old[k], new[k])
    elif isinstance(old, (list, tuple)):
        for a, b in zip(old, new):
            _copy_tree(a, b)


class FusedTransGoInference:
    def __init__(self, model, dtype=torch.bfloat16, attention="tcgen05", head_conv="mma", stem_conv="mma"):
        """attention: "tcgen05" (default: the one-kernel block wherever it applies) or "mma" (GEMM + tg_net_attention everywhere;
        kept as the cross-check and for timing comparisons).  head_conv / stem_conv: "mma" (default: tg_net_conv3x3_head / tg_net_stem on
        9x9) or "library"."""
        self.model, self.dtype, self.attention, self.head_conv, self.stem_conv = model, dtype, attention, head_conv, stem_conv
        self.lib = _lib.lib()
        self._w = None
        self.refresh()

    # ------------------------------------------------------------------ weight folding
    def refresh(self):
        """(Re-)fold the module's current weights.  After the first call the folded tensors are updated IN PLACE, so CUDA
        graphs that captured this forward (the evaluator's own, or a whole search wave) keep seeing the new weights."""
        new = self._fold_all()
        if getattr(self, "_w", None) is None:
            self._w = new
            for k, v in new.items():
                setattr(self, k, v)
            self._bufs = {}
        else:
            _copy_tree(self._w, new)

    def _fold_all(self):
        m, dt = self.model.main_network, self.dtype
        self.n = m.board_width
        self.P = self.n * self.n
        self.device = next(m.parameters()).device
        w = {}
        w["stem"] = _conv_bn(m.conv1.conv[0], m.conv1.conv[1], dt, pad_in=16)
        ws, bs = (t.float() for t in _conv_bn(m.conv1.conv[0], m.conv1.conv[1], torch.float32))      # the same, packed for tg_net_stem:
        cin = ws.shape[1]                                                                              # [128][tap][16 input channels]
        w["stem_k"] = (F.pad(ws.permute(0, 2, 3, 1).reshape(ws.shape[0], 9, cin), (0, 16 - cin)).to(dt).contiguous(), bs.contiguous()) \
            if cin <= 12 else None
        self.cin = cin
        blocks = []
        i = 2
        while hasattr(m, "res_conv%d" % i):
            blk = getattr(m, "res_conv%d" % i)
            if isinstance(blk, _PreActResidual):
                s1, t1 = _fold(blk.batchnormlize_1)
                blocks.append(("R", dict(s1=s1.contiguous(), t1=t1.contiguous(),
                                         c1=_conv_bn(blk.conv_1, blk.batchnormlize_2, dt), c2=_conv_bn(blk.conv_2, None, dt),
                                         c2b=blk.conv_2.bias.detach().float().contiguous())))
            else:
                blocks.append(("A", self._attn(blk)))
            i += 1
        w["blocks"] = blocks
        w["end"] = tuple(t.contiguous() for t in _fold(m.bn_res_end))
        P = self.P

        def nhwc_fc(fc, ch):                   # the reference flattens NCHW (c*P + p); our activations flatten as p*ch + c
            wt = fc.weight.detach().float().view(fc.out_features, ch, P).permute(0, 2, 1).reshape(fc.out_features, P * ch)
            return wt.to(dt).contiguous(), fc.bias.detach().to(dt).contiguous()
        vc = _conv_bn(m.conv_val_own.conv[0], m.conv_val_own.conv[1], dt)
        w["val_conv"] = (F.pad(vc[0], (0, 0, 0, 0, 0, 0, 0, 6)).contiguous(memory_format=torch.channels_last),
                         F.pad(vc[1], (0, 6)))                                  # 2 -> 8 output channels (tensor-core friendly)
        w["val_head"] = _conv_head(m.conv_val_own.conv[0], m.conv_val_own.conv[1], dt)       # the same two convolutions packed for
        w["act_head"] = _conv_head(m.conv_act.conv[0], m.conv_act.conv[1], dt)               # tg_net_conv3x3_head (9x9 boards)
        w["fc_val_own"] = nhwc_fc(m.fc_val_own, 2)
        w["fc_val"] = (m.fc_val.weight.detach().to(dt).contiguous(), m.fc_val.bias.detach().to(dt).contiguous())
        w["fc_own"] = (m.fc_own.weight.detach().to(dt).contiguous(), m.fc_own.bias.detach().to(dt).contiguous())
        # value and ownership share their input: one GEMM with the rows stacked, output [1 + P, B] so that the value row is contiguous
        w["fc_vo"] = (torch.cat([m.fc_val.weight.detach(), m.fc_own.weight.detach()], 0).to(dt).contiguous(),
                      torch.cat([m.fc_val.bias.detach(), m.fc_own.bias.detach()], 0).to(dt).reshape(-1, 1).contiguous())
        w["act_attn"] = self._attn(m.attention_act)
        ac = _conv_bn(m.conv_act.conv[0], m.conv_act.conv[1], dt)
        w["act_conv"] = (F.pad(ac[0], (0, 0, 0, 0, 0, 0, 0, 4)).contiguous(memory_format=torch.channels_last),
                         F.pad(ac[1], (0, 4)))                                  # 4 -> 8 output channels
        w["fc_act"] = nhwc_fc(m.fc_act, 4)
        return w

    def _attn(self, blk):
        """q | k | v as one GEMM.  gamma (model.py:312) is folded into the value projection -- gamma * sum_i att * v_i =
        sum_i att * (gamma * v_i) -- so that no host scalar is baked into captured graphs."""
        dt = self.dtype
        wq, wk, wv = (c.weight.detach().float().flatten(1) for c in (blk.query_conv, blk.key_conv, blk.value_conv))
        bq, bk, bv = (c.bias.detach().float() for c in (blk.query_conv, blk.key_conv, blk.value_conv))
        gamma = blk.gamma.detach().float().reshape(1, 1)
        s, t = _fold(blk.bn)
        return dict(w=torch.cat([wq, wk, wv * gamma], 0).to(dt).contiguous(), b=torch.cat([bq, bk, bv * gamma.reshape(1)], 0).to(dt).contiguous(),
                    dq=wq.shape[0], gamma=1.0, s=s.contiguous(), t=t.contiguous())

    # ------------------------------------------------------------------ pieces
    def _junction(self, x, y, alpha, scale, shift, want_sum, want_act, ybias=None):
        """x (+ alpha*(y + ybias)) -> (sum, relu(sum*scale+shift)); x, y: channels_last [B,C,n,n] bf16 (NHWC memory)"""
        B, C = x.shape[0], x.shape[1]
        if self.dtype == torch.float32 and x.is_contiguous(memory_format=torch.channels_last) and (y is None or y.is_contiguous(memory_format=torch.channels_last)):
            s_out = torch.empty_like(x) if want_sum else None            # the float32 twin of the junction kernel
            a_out = torch.empty_like(x) if want_act else None
            _lib.check(self.lib.tg_net_add_bn_relu_f32(_lib.ptr(x), _lib.ptr(y), float(alpha), _lib.ptr(ybias), _lib.ptr(scale), _lib.ptr(shift),
                                                       _lib.ptr(s_out), _lib.ptr(a_out), B * self.P, C, _lib.stream_ptr(device=x.device)),
                       "tg_net_add_bn_relu_f32")
            return s_out, a_out
        if self.dtype != torch.bfloat16:
            # other copies of the same schedule (float16, or float32 tensors that are not NHWC): library elementwise ops
            s = x if y is None else x + alpha * (y if ybias is None else y + ybias.to(y.dtype).view(1, -1, 1, 1))
            a = torch.relu(s * scale.to(s.dtype).view(1, -1, 1, 1) + shift.to(s.dtype).view(1, -1, 1, 1)) if want_act else None
            return (s if want_sum else None), a
        rows = B * self.P
        s_out = torch.empty_like(x) if want_sum else None
        a_out = torch.empty_like(x) if want_act else None
        _lib.check(self.lib.tg_net_add_bn_relu(_lib.ptr(x), _lib.ptr(y), float(alpha), _lib.ptr(ybias), _lib.ptr(scale), _lib.ptr(shift),
                                               _lib.ptr(s_out), _lib.ptr(a_out), rows, C, _lib.stream_ptr(device=x.device)),
                   "tg_net_add_bn_relu")
        return s_out, a_out

    def _attention(self, x, p, nxt=None):
        """model.py:299-314 on an NHWC tensor: returns relu(bn(gamma * (V . softmax(Q^T K)) + x)) and, when the next block is a
        residual block (nxt = its folded first BatchNorm (scale, shift)) and the one-kernel path runs, that block's opening
        activation relu(bn1(out)) as a second output (None otherwise: the caller then runs the junction kernel)"""
        B, C = x.shape[0], x.shape[1]
        dq = p["dq"]
        if self.attention == "tcgen05" and self.P <= 96 and dq == 32 and C == 128 and self.dtype == torch.bfloat16:
            # projection, scores, softmax, att^T . v and the junction in ONE tcgen05 kernel (csrc/tg_net_tc.cu: k_attention_tc)
            act = torch.empty_like(x)
            act2 = torch.empty_like(x) if nxt is not None else None
            _lib.check(self.lib.tg_net_attention_fused2(_lib.ptr(x), _lib.ptr(p["w"]), _lib.ptr(p["b"]), p["gamma"], _lib.ptr(p["s"]),
                                                        _lib.ptr(p["t"]), _lib.ptr(act), _lib.ptr(nxt[0]) if nxt else None,
                                                        _lib.ptr(nxt[1]) if nxt else None, _lib.ptr(act2), B, self.P, dq, C, None,
                                                        _lib.stream_ptr(device=x.device)), "tg_net_attention_fused")
            return act, act2
        x2 = x.permute(0, 2, 3, 1).reshape(B, self.P, C)                       # free view of channels_last memory
        qkv = F.linear(x2, p["w"], p["b"])
        if self.P <= 368 and dq == 32 and C == 128 and self.dtype == torch.bfloat16:
            # scores, softmax, att^T . v and the junction in ONE hand-written kernel (csrc/tg_net.cu: k_attention)
            act = torch.empty_like(x)
            act2 = torch.empty_like(x) if nxt is not None else None
            _lib.check(self.lib.tg_net_attention2(_lib.ptr(qkv), _lib.ptr(x), p["gamma"], _lib.ptr(p["s"]), _lib.ptr(p["t"]), _lib.ptr(act),
                                                  _lib.ptr(nxt[0]) if nxt else None, _lib.ptr(nxt[1]) if nxt else None, _lib.ptr(act2),
                                                  B, self.P, dq, C, _lib.stream_ptr(device=x.device)), "tg_net_attention")
            return act, act2
        q, k, v = qkv[..., :dq], qkv[..., dq:2 * dq], qkv[..., 2 * dq:]
        att = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)           # [B, P, P], rows = query position
        o = torch.bmm(att.transpose(1, 2), v)                                  # out[:, j] = sum_i att[i, j] * v[i]
        o = o.reshape(B, self.n, self.n, C).permute(0, 3, 1, 2)                # NHWC memory, NCHW view
        _, act = self._junction(x, o, p["gamma"], p["s"], p["t"], False, True)
        return act, None

    def _head_conv(self, x, packed, lib_w, nout):
        """relu(bn(conv3x3(x))) with `nout` output channels, flattened per position in (point, channel) order: [B, P * nout]"""
        B = x.shape[0]
        if self.head_conv == "mma" and self.n == 9 and x.shape[1] == 128 and self.dtype == torch.bfloat16:
            out = torch.empty([B, self.P * nout], dtype=self.dtype, device=x.device)
            _lib.check(self.lib.tg_net_conv3x3_head(_lib.ptr(x), _lib.ptr(packed[0]), _lib.ptr(packed[1]), _lib.ptr(out), B, self.n, 128, nout,
                                                    _lib.stream_ptr(device=x.device)), "tg_net_conv3x3_head")
            return out
        h = torch.cudnn_convolution_relu(x, lib_w[0], lib_w[1], [1, 1], [1, 1], [1, 1], 1)[:, :nout]
        return h.permute(0, 2, 3, 1).reshape(B, self.P * nout)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, obs):
        """obs float32 [B, 10, n, n] -> (policy float32 [B, A], value float32 [B], ownership float32 [B, P])"""
        B = obs.shape[0]
        key = (B, obs.device)
        if key not in self._bufs:
            self._bufs[key] = torch.zeros([B, 16, self.n, self.n], dtype=self.dtype, device=obs.device).contiguous(memory_format=torch.channels_last)
        conv_relu = torch.cudnn_convolution_relu
        one, pad = [1, 1], [1, 1]
        pending = None                                                         # activation already prepared for the next R block
        nb = len(self.blocks)
        if (self.stem_conv == "mma" and self.n == 9 and self.dtype == torch.bfloat16 and self.stem_k is not None and obs.dtype == torch.float32
                and obs.is_contiguous() and self.stem[0].shape[0] == 128):
            # planes -> NHWC bf16 trunk input in one hand-written kernel, with the first residual block's opening activation
            x = torch.empty([B, 128, self.n, self.n], dtype=self.dtype, device=obs.device).contiguous(memory_format=torch.channels_last)
            first = self.blocks[0] if nb and self.blocks[0][0] == "R" else None
            pending = torch.empty_like(x) if first is not None else None
            _lib.check(self.lib.tg_net_stem(_lib.ptr(obs), self.cin, _lib.ptr(self.stem_k[0]), _lib.ptr(self.stem_k[1]),
                                            _lib.ptr(first[1]["s1"]) if first else None, _lib.ptr(first[1]["t1"]) if first else None,
                                            _lib.ptr(x), _lib.ptr(pending), B, self.n, 128, _lib.stream_ptr(device=obs.device)), "tg_net_stem")
        else:
            x16 = self._bufs[key]
            x16[:, :obs.shape[1]] = obs
            x = conv_relu(x16, self.stem[0], self.stem[1], one, pad, one, 1)   # stem: conv + folded BN + ReLU
        for i, (kind, p) in enumerate(self.blocks):
            if kind == "R":
                a1 = pending if pending is not None else self._junction(x, None, 1.0, p["s1"], p["t1"], False, True)[1]
                pending = None
                h = conv_relu(a1, p["c1"][0], p["c1"][1], one, pad, one, 1)    # conv_1 + folded batchnormlize_2 + ReLU
                y = F.conv2d(h, p["c2"][0], None, padding=1)                  # bias-free: conv_2's bias rides in the junction
                b2 = p["c2b"]
                nxt = self.blocks[i + 1] if i + 1 < nb else None
                if nxt is not None and nxt[0] == "R":                          # x + y and the next block's BN1 + ReLU in one pass
                    x, pending = self._junction(x, y, 1.0, nxt[1]["s1"], nxt[1]["t1"], True, True, b2)
                elif nxt is None:                                              # last block: x + y, bn_res_end, ReLU
                    x = self._junction(x, y, 1.0, self.end[0], self.end[1], False, True, b2)[1]
                else:
                    x = self._junction(x, y, 1.0, None, None, True, False, b2)[0]
            else:
                nxt = self.blocks[i + 1] if i + 1 < nb else None
                x, pending = self._attention(x, p, (nxt[1]["s1"], nxt[1]["t1"]) if nxt is not None and nxt[0] == "R" else None)
        # value / ownership head (model.py:93-99)
        hv = self._head_conv(x, self.val_head, self.val_conv, 2)
        hv = F.relu(F.linear(hv, *self.fc_val_own))
        vo = torch.tanh(torch.addmm(self.fc_vo[1], self.fc_vo[0], hv.t()).float())     # [1 + P, B]: tanh of (W . h^T + b), like the two
        value, own = vo[0], vo[1:].t()                                                 # separate F.linear calls (same bf16 rounding)
        # policy head (model.py:103-111)
        xa, _ = self._attention(x, self.act_attn)
        ha = self._head_conv(xa, self.act_head, self.act_conv, 4)
        policy = torch.softmax(F.linear(ha, *self.fc_act), -1, dtype=torch.float32)     # the cast happens inside the softmax kernel
        return policy, value, own


def make_fused_evaluator(model, dtype=torch.bfloat16, cuda_graph=True, attention="tcgen05", head_conv="mma", stem_conv="mma"):
    """evaluator(obs) -> (policy, value) for BatchedMCTS, forward captured once per static input tensor as a CUDA graph.
    cuda_graph=False: plain launches (use it when the caller captures a larger graph, e.g. BatchedMCTS(graph_waves=True));
    cuda_graph="roots": plain launches while the current stream is being captured, own graphs otherwise (the search's wave
    batch is captured by the engine, the per-move root batch runs outside it)."""
    net = FusedTransGoInference(model, dtype, attention=attention, head_conv=head_conv, stem_conv=stem_conv)
    graphs = {}

    def evaluator(obs):
        if not cuda_graph or (cuda_graph == "roots" and torch.cuda.is_current_stream_capturing()):
            p, v, _ = net(obs)
            return p, v
        key = (obs.data_ptr(), tuple(obs.shape))
        if key not in graphs:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    net(obs)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = net(obs)
            graphs[key] = (g, out)
        g, out = graphs[key]
        g.replay()
        return out[0], out[1]
    evaluator.net = net
    evaluator.refresh = net.refresh                       # in place: captured graphs stay valid
    return evaluator
