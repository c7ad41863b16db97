"""The policy / value / ownership net the search consumes (reference: model.py TransGoNetwork / MainNetwork, :11-114).

north_star keeps this net in PyTorch (tensor cores are used only here, through cuDNN / cuBLAS); it is re-stated so that
the engine is self-contained on machines without the reference, with the SAME parameter names and shapes, so a reference
checkpoint ({"weights": state_dict}, transgo.py:171-184) loads with ``set_weights`` unchanged:

  main_network.conv1.conv.{0,1}            3x3 conv 10->F + BN + ReLU                      (model.py:49, :316-324)
  main_network.res_conv{2,4,5,6,8..11,13}  pre-activation residual block, 2 x 3x3 conv    (model.py:208-262)
  main_network.res_conv{3,7,12}            single-head self-attention + BN + ReLU         (model.py:286-314)
  main_network.bn_res_end                  BN + ReLU
  value head   conv_val_own (3x3 F->2) -> fc_val_own (2P->64) -> fc_val tanh(1), fc_own tanh(P)
  policy head  attention_act -> conv_act (3x3 F->4) -> fc_act (4P->P+1) -> softmax
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _ConvBnRelu(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.conv(x)


class _PreActResidual(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.batchnormlize_1 = nn.BatchNorm2d(ch)
        self.conv_shortcut = nn.Conv2d(ch, ch, 1)          # present in reference checkpoints; unused when in == out
        self.conv_1 = nn.Conv2d(ch, ch, 3, 1, 1)
        self.conv_2 = nn.Conv2d(ch, ch, 3, 1, 1)
        self.batchnormlize_2 = nn.BatchNorm2d(ch)

    def forward(self, x):
        y = self.conv_1(F.relu(self.batchnormlize_1(x)))
        y = self.conv_2(F.relu(self.batchnormlize_2(y)))
        return x + y


class _SpatialAttention(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.query_conv = nn.Conv2d(ch, ch // 4, 1)
        self.key_conv = nn.Conv2d(ch, ch // 4, 1)
        self.value_conv = nn.Conv2d(ch, ch, 1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.bn = nn.BatchNorm2d(ch)

    def forward(self, x):
        b, c, h, w = x.shape
        q = self.query_conv(x).flatten(2).transpose(1, 2)          # B x P x C/4
        k = self.key_conv(x).flatten(2)                            # B x C/4 x P
        att = torch.softmax(torch.bmm(q, k), dim=-1)               # B x P x P
        v = self.value_conv(x).flatten(2)                          # B x C x P
        out = torch.bmm(v, att).view(b, c, h, w)
        return F.relu(self.bn(self.gamma * out + x))


_TRUNK = ("R", "A", "R", "R", "R", "A", "R", "R", "R", "R", "A", "R")      # res_conv2 .. res_conv13 (model.py:50-61)


class MainNetwork(nn.Module):
    def __init__(self, board_width, board_height, input_dim, num_features):
        super().__init__()
        self.board_width, self.board_height = board_width, board_height
        p = board_width * board_height
        self.conv1 = _ConvBnRelu(input_dim, num_features)
        for i, kind in enumerate(_TRUNK):
            setattr(self, "res_conv%d" % (i + 2), _PreActResidual(num_features) if kind == "R" else _SpatialAttention(num_features))
        self.bn_res_end = nn.BatchNorm2d(num_features)
        self.conv_val_own = _ConvBnRelu(num_features, 2)
        self.fc_val_own = nn.Linear(2 * p, 64)
        self.fc_val = nn.Linear(64, 1)
        self.fc_own = nn.Linear(64, p)
        self.attention_act = _SpatialAttention(num_features)
        self.conv_act = _ConvBnRelu(num_features, 4)
        self.fc_act = nn.Linear(4 * p, p + 1)

    def forward(self, x):
        x = self.conv1(x)
        for i in range(len(_TRUNK)):
            x = getattr(self, "res_conv%d" % (i + 2))(x)
        x = F.relu(self.bn_res_end(x))
        h = F.relu(self.fc_val_own(self.conv_val_own(x).flatten(1)))
        value, own = torch.tanh(self.fc_val(h)), torch.tanh(self.fc_own(h))
        policy = torch.softmax(self.fc_act(self.conv_act(self.attention_act(x)).flatten(1)), -1)
        return policy, value, own


class TransGoNetwork(nn.Module):
    """model.py:11-27 surface: main_prediction / get_weights / set_weights."""

    def __init__(self, config=None, board_size=None, input_dim=10, num_features=128):
        super().__init__()
        n = board_size or getattr(config, "board_size", 9)
        self.main_network = MainNetwork(n, n, getattr(config, "input_dim", input_dim), getattr(config, "num_features", num_features))

    def main_prediction(self, state):
        return self.main_network(state)

    def get_weights(self):
        return {k: v.cpu() for k, v in self.state_dict().items()}

    def set_weights(self, weights):
        self.load_state_dict(weights)


def make_evaluator(model, dtype=None, channels_last=False, cuda_graph=False):
    """obs [B,10,n,n] float32 -> (policy float32 [B,A], value float32 [B]) under no_grad.

    dtype None / float32: the model as it is (parity runs).  dtype bfloat16 / float16: an inference COPY of the model with
    weights converted once (and channels_last when asked) -- 1.4x faster than autocast at the wave batch on B200 because
    no per-call casts remain.  cuda_graph: capture the forward once per input shape and replay it (the search always
    calls with the same two static input tensors: the leaf batch and the root batch)."""
    import copy
    model.eval()
    low = dtype is not None and dtype != torch.float32
    run_model = model
    if low:
        run_model = copy.deepcopy(model).to(dtype)
        if channels_last:
            run_model = run_model.to(memory_format=torch.channels_last)
        run_model.eval()
    graphs = {}

    def forward(obs):
        x = obs.to(dtype) if low else obs
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        p, v, _ = run_model.main_prediction(x)
        return p.float(), v.float().reshape(-1)

    def evaluator(obs):
        with torch.no_grad():
            if not cuda_graph:
                return forward(obs)
            key = (obs.data_ptr(), tuple(obs.shape))
            if key not in graphs:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(3):
                        forward(obs)
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = forward(obs)
                graphs[key] = (g, out)
            g, out = graphs[key]
            g.replay()
            return out
    evaluator.refresh = lambda: run_model.load_state_dict({k: v.to(dtype) if v.is_floating_point() else v for k, v in model.state_dict().items()}) if low else None
    return evaluator
