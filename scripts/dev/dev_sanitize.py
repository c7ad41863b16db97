"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python scripts/dev_sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from transgo_b200 import _lib
from transgo_b200.environment import BatchedGoEnv
from transgo_b200.mcts import BatchedMCTS, _declare
L = _lib.lib(); _declare(L)
for n, G in ((9, 70), (19, 9)):
    for tile in (0, 32, 16, 8, 1, -1, 8 | 64, 4 | 64):
        env = BatchedGoEnv(G, board_size=n, tile=tile)
        trace = torch.zeros([env.max_step + 1, G], dtype=torch.int16, device="cuda")
        cnt = torch.zeros([1], dtype=torch.int64, device="cuda")
        env.playout(min(env.max_step + 1, 150), seed=3, trace=trace, counter=cnt, no_eye=(tile == 8))
        env.step(torch.full([G], n * n, dtype=torch.int16, device="cuda"))
        env.legal_mask(); env.legal_mask(no_eye=True); env.encode(); env.score(want_territory=True)
    torch.cuda.synchronize()
    print("board kernels ok", n)
for n, G, sims in ((9, 12, 40), (19, 3, 16)):
    A = n * n + 1
    pol = {G * 4: torch.empty([G * 4, A], device="cuda"), G: torch.empty([G, A], device="cuda")}
    val = {G * 4: torch.empty([G * 4], device="cuda"), G: torch.empty([G], device="cuda")}
    wave = torch.zeros([1], dtype=torch.int64, device="cuda")

    def synth(obs, n=n, pol=pol, val=val, wave=wave):
        B = obs.shape[0]
        _lib.check(L.tg_synth_provider_dev(n, B, 5, _lib.ptr(wave), _lib.ptr(pol[B]), _lib.ptr(val[B]), _lib.stream_ptr()))
        return pol[B], val[B]
    eng = BatchedMCTS(synth, G, board_size=n, num_simulation=sims, tie_seed=2, max_step=12)
    for _ in range(6):
        a, _, _ = eng.get_action_probs(True)
        eng.update_with_action(a)
    eng.search_stats(); eng.outcome()
    torch.cuda.synchronize()
    print("mcts kernels ok", n)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for B, P in ((3, 81), (2, 361), (2, 97)):
    qkv = torch.randn([B, P, 192], device="cuda", generator=g).to(torch.bfloat16)
    x = torch.randn([B, P, 128], device="cuda", generator=g).to(torch.bfloat16)
    sc = torch.rand([128], device="cuda", generator=g); sh = torch.rand([128], device="cuda", generator=g)
    out = torch.empty_like(x)
    _lib.check(L.tg_net_attention(_lib.ptr(qkv), _lib.ptr(x), 0.5, _lib.ptr(sc), _lib.ptr(sh), _lib.ptr(out), B, P, 32, 128, _lib.stream_ptr()))
    s_out = torch.empty_like(x); a_out = torch.empty_like(x)
    _lib.check(L.tg_net_add_bn_relu(_lib.ptr(x), _lib.ptr(out), 1.0, _lib.ptr(sh), _lib.ptr(sc), _lib.ptr(sh), _lib.ptr(s_out), _lib.ptr(a_out), B * P, 128, _lib.stream_ptr()))
torch.cuda.synchronize()
print("net kernels ok")
