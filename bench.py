#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the batched self-play hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[1]): 9x9 env-only batched random playouts, 4096 concurrent
games per GPU from the empty board to termination.  One *step* = one such batch: reset + playout kernel (every ply:
group analysis -> legal mask -> hashed move draw -> capture/ko/step) + Tromp-Taylor scoring.
  value      board steps (plies) per second, whole job, game states resident in HBM, CUDA-event timed
  e2e        same metric through the C ABI with HOST buffers, all transfers inside the timed region; two routes are timed and the
             faster one is reported (both under e2e.routes): "copies" = pinned initial states H2D, kernels, final states + scores +
             the per-ply action trace D2H as separate copy operations; "zero_copy" = ONE launch (tg_env_playout_io) whose kernel
             reads the pinned start positions and writes final positions, scores and trace into pinned host memory across PCIe itself
  roofline   k_playout against the measured HBM peak with SURVEY 8d's 115 algorithmic bytes per ply
  cpu_baseline  the UNMODIFIED reference engine (oracle/_ref/go_env9.so, compiled from /root/reference) driven through
             ctypes exactly like GoEnv/environment.py (getLegalAction + Step per ply) on every host core, same move rule
  mcts       second half of the metric ("MCTS sims/sec"): BASELINE configs[2] (1024 games x 800 simulations per move),
             tree kernels alone with a device-side synthetic provider, and end to end with the random-init TransGo net
Multi-GPU: games shard by id (rank r owns games [r*G, (r+1)*G)); no collective on the data path; scaling is weak.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BYTES_PER_PLY = 115          # read S(48) + action(2); write S(48) + legal mask(16) + done(1)
SEED = 0x7261
GAMES = 4096
MCTS_GAMES, MCTS_SIMS = 1024, 800


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
    """One host core: reference engine (or, if it is not built, the oracle port) playing whole games with the C2 move rule."""
    wid, seconds, game_base = args
    from oracle.go_oracle import OracleEnv, RefEnv, have_ref
    from helpers import playout_action
    kind = "reference" if have_ref(9) else "port"
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    env = RefEnv(9) if kind == "reference" else OracleEnv(9)
    plies, games = 0, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        s, _ = env.reset()
        done = False
        g = game_base + games
        while not done:
            acts = env.getLegalAction(s)                       # environment.py:121-129 (ctypes + python list filter)
            pts = [a for a in acts if a != 81]
            s, done = env.step(s, playout_action(SEED, g, env.getStep(s), pts, 81))
            plies += 1
        env.getScore(s)
        games += 1
    return plies, games, time.perf_counter() - t0, kind


def cpu_playout_baseline(seconds_per_core=1.5, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(w, seconds_per_core, w * 100000) for w in range(cores)])
    plies = sum(r[0] for r in res)
    wall = max(r[2] for r in res)
    return dict(value=plies / wall, unit="board steps/s", cores=cores, kind=res[0][3],
                sample="%d whole 9x9 games (%d plies) of the C2 move rule, %d processes x %.1f s, via ctypes like environment.py"
                       % (sum(r[1] for r in res), plies, cores, seconds_per_core)), (plies, sum(r[1] for r in res), wall)


def _mcts_kind():
    """What the CPU MCTS arm runs: the reference's own WP_MCTS + GoEnv + go_env.so (needs /root/reference: build container
    only), else the pinned port of the search (oracle/mcts_oracle.py) over the UNMODIFIED reference engine
    (oracle/_ref/go_env9.so travels to the GPU box), else the port over the C restatement."""
    from oracle import ref_mcts_harness
    from oracle.go_oracle import have_ref
    if ref_mcts_harness.available():
        return "reference"
    return "port-search+reference-engine" if have_ref(9) else "port"


def _cpu_mcts_worker(args):
    wid, sims = args[0], args[1]
    net_device = args[2] if len(args) > 2 else "cpu"
    import torch
    torch.set_num_threads(1)
    import numpy as np
    kind = _mcts_kind()
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    if kind == "reference":
        # self_play.py:575-875 + GoEnv/environment.py:32-198 + model.py, imported where they lie; ray is stubbed (not installed)
        from oracle import ref_mcts_harness as H
        self_play, configure = H._import_reference()
        import model as ref_model
        cfg = configure.Config()
        cfg.device = torch.device(net_device)
        cfg.num_simulation = sims
        cwd = os.getcwd()
        os.chdir(H.REFERENCE)                                  # environment.py:42 loads ./GoEnv/go_env.so relative to the CWD
        try:
            env = self_play.GoEnv(cfg)
        finally:
            os.chdir(cwd)
        torch.manual_seed(0)
        net = ref_model.TransGoNetwork(cfg).to(cfg.device).eval()
        with torch.no_grad():
            m = self_play.WP_MCTS(cfg, env, net, None)         # constructor expands the root (one net call)
            t0 = time.perf_counter()
            m.get_action_probs(True)
            dt = time.perf_counter() - t0
        return m.root.total_visit_count, dt, kind
    from oracle.go_oracle import OracleEnv, RefEnv
    from oracle.mcts_oracle import OracleMCTS, MCTSConfig
    from oracle import provider
    from transgo_b200.net import TransGoNetwork
    torch.manual_seed(0)
    net = TransGoNetwork(board_size=9).eval().to(net_device)

    def ev(obs):                                               # self_play.py:777-793: H2D, forward, three D2H per wave
        with torch.no_grad():
            p, v, _ = net.main_prediction(torch.from_numpy(np.stack(obs)).to(net_device))
        return p.cpu().numpy(), v.cpu().numpy()
    if net_device != "cpu":
        ev([np.zeros((10, 9, 9), np.float32)] * 4)             # CUDA context + cuDNN autotune outside the timed move
    cnt = [0]

    def tie(n):
        i = provider.tie_index(1, wid, cnt[0], n); cnt[0] += 1
        return i
    env = RefEnv(9) if kind == "port-search+reference-engine" else OracleEnv(9)
    m = OracleMCTS(env, ev, MCTSConfig(num_simulation=sims), tie_fn=tie,
                   noise_fn=lambda mv, k: provider.dirichlet(1, wid, mv, k), u_fn=lambda mv: provider.uniform(1, wid, mv))
    t0 = time.perf_counter()
    m.get_action_probs()
    return m.sims, time.perf_counter() - t0, kind


def cpu_mcts_baseline(cores=None, sims=210, net_device="cpu"):
    """BASELINE configs[0] on the host: ONE game per process at the reference defaults (configure.py:25-33: 210 simulations,
    4 leaves per net call), one move after the root expansion, `cores` processes like the reference's actors
    (transgo.py:92-96).  kind "reference" = self_play.py's WP_MCTS itself; elsewhere the pinned port of the search over the
    reference engine.  net_device "cpu": one CPU thread per process; "cuda:0": the reference as shipped (configure.py:66)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_mcts_worker, [(w, sims, net_device) for w in range(cores)])
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    where = "net on 1 CPU thread each" if net_device == "cpu" else "python tree search on the host, net forward on %s as the reference ships" % net_device
    what = {"reference": "the reference's WP_MCTS + GoEnv + go_env.so + model.py (imported from /root/reference)",
            "port-search+reference-engine": "oracle/mcts_oracle.py (pinned bit-exact to WP_MCTS) over the unmodified reference engine oracle/_ref/go_env9.so",
            "port": "oracle/mcts_oracle.py over oracle/go_rules.c"}[res[0][2]]
    return dict(value=total / wall, unit="MCTS sims/s", cores=cores, kind=res[0][2],
                sample="%d processes x one move of %d simulations (4 leaves per net call, %s); %s" % (cores, sims, where, what))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """NVML samples DURING the timed regions (B200_PROFILING.md 'clocks line')."""

    def __init__(self, index):
        self.samples, self.reasons, self._stop, self.ok = [], set(), threading.Event(), False
        self.active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"} \
            if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else \
            {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
             nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.005)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()

    def summary(self):
        if not self.ok or not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=getattr(self, "max", None), reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(self.samples), sm_max_mhz=self.max, reasons=sorted(self.reasons), samples=len(self.samples))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    cpu = None
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_mcts = cpu_mcts_gpu_net = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = cpu_playout_baseline()                        # before CUDA is initialised (fork-safe)
        if not args.no_mcts:
            cpu_mcts = cpu_mcts_baseline(sims=64 if args.quick else 210)
            try:                                               # children create their own CUDA contexts; the parent has none yet
                cpu_mcts_gpu_net = cpu_mcts_baseline(cores=6, sims=210, net_device="cuda:%d" % local)
            except Exception as e:  # pragma: no cover
                cpu_mcts_gpu_net = {"unavailable": str(e)[:200]}
    import torch
    import torch.distributed as dist
    from transgo_b200 import _lib
    from transgo_b200.environment import BatchedGoEnv
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    K, W, G = args.steps, max(args.warmup, 3), args.games
    env = BatchedGoEnv(G, board_size=9, tile=args.tile, device=dev)
    plies_max = env.max_step + 1
    counter = torch.zeros([1], dtype=torch.int64, device=dev)
    flush = torch.empty([64 << 20], dtype=torch.int32, device=dev)      # 256 MB > 126 MB L2
    sampler = ClockSampler(local)
    sampler.start()
    launches = [0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    d_scores = torch.empty([G], dtype=torch.float32, device=dev)

    def device_step(step_idx):
        # reset + playout + final score: ONE launch (tg_env_playout_scored)
        env.playout(plies_max, seed=SEED + step_idx, game_id0=rank * G, counter=counter, from_reset=True, score_out=d_scores)
        launches[0] += 1

    # ---- device-resident throughput
    for i in range(W):
        device_step(1000 + i)
    counter.zero_()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    launches[0] = 0
    sampler.active.set()
    for i in range(K):
        flush.zero_()                                          # L2 flush between timed iterations (outside the events)
        evs[i][0].record()
        kevs[i][0].record()
        env.playout(plies_max, seed=SEED + i, game_id0=rank * G, counter=counter, from_reset=True, score_out=d_scores)
        kevs[i][1].record()
        evs[i][1].record()
        launches[0] += 1
    barrier()
    sampler.active.clear()
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    ms_kernel = sum(a.elapsed_time(b) for a, b in kevs)
    plies = int(counter.item())
    t = torch.tensor([ms_total, ms_kernel], dtype=torch.float64, device=dev)
    p = torch.tensor([plies], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(p, op=dist.ReduceOp.SUM)
    ms_total, ms_kernel, plies_all = float(t[0]), float(t[1]), float(p[0])
    value = plies_all / (ms_total * 1e-3)

    # ---- end to end through the C ABI with host buffers
    SQ = env.SQ
    h_states = torch.zeros([SQ, G, 4], dtype=torch.int32).pin_memory()
    fresh = BatchedGoEnv(G, board_size=9, device=dev)
    h_states.copy_(fresh.states)                               # empty boards, prepared on the host side once
    h_out_states = torch.zeros([SQ, G, 4], dtype=torch.int32).pin_memory()
    h_scores = torch.zeros([G], dtype=torch.float32).pin_memory()
    h_trace = torch.zeros([plies_max, G], dtype=torch.int16).pin_memory()
    d_trace = torch.zeros([plies_max, G], dtype=torch.int16, device=dev)
    h2d = h_states.numel() * 4
    d2h = h_out_states.numel() * 4 + h_scores.numel() * 4 + h_trace.numel() * 2

    def e2e_step(i):
        env.states.copy_(h_states, non_blocking=True)
        env.playout(plies_max, seed=SEED + i, game_id0=rank * G, trace=d_trace, counter=counter, score_out=d_scores)
        h_out_states.copy_(env.states, non_blocking=True)
        h_scores.copy_(d_scores, non_blocking=True)
        h_trace.copy_(d_trace, non_blocking=True)
    # the same bytes without separate copy operations: the kernel reads the pinned host states and writes the final states, the
    # scores and the move trace into pinned host memory itself (zero-copy over PCIe; every buffer still crosses the bus once)
    h_zc_states = torch.zeros([SQ, G, 4], dtype=torch.int32).pin_memory()
    h_zc_scores = torch.zeros([G], dtype=torch.float32).pin_memory()
    h_zc_trace = torch.zeros([plies_max, G], dtype=torch.int16).pin_memory()

    def time_e2e(step_fn, prep=None):
        for i in range(W):
            if prep:
                prep()
            step_fn(2000 + i)
        torch.cuda.synchronize()
        counter.zero_()
        ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        sampler.active.set()
        for i in range(K):
            flush.zero_()
            if prep:
                torch.cuda.synchronize()                       # the previous launch has finished with the host buffer
            ee[i][0].record()
            if prep:
                prep()                                         # host-side input preparation INSIDE the timed region
            step_fn(i)
            ee[i][1].record()
        barrier()
        sampler.active.clear()
        ms = sum(a.elapsed_time(b) for a, b in ee)
        n = int(counter.item())
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        p = torch.tensor([n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(p, op=dist.ReduceOp.SUM)
        return float(p[0]) / (float(t[0]) * 1e-3), float(t[0])

    e2e_copies_value, ms_e2e_copies = time_e2e(e2e_step)
    assert bool((h_scores * 2 % 2 == 1).all())                 # host really received final area scores

    def zc_launch(i):
        # start positions are READ from h_states and the final ones written to h_zc_states (tg_env_playout_io): nothing to refill
        env.playout(plies_max, seed=SEED + i, game_id0=rank * G, trace=h_zc_trace, counter=counter, score_out=h_zc_scores,
                    states=h_zc_states, states_in=h_states)
    e2e_zc_value, ms_e2e_zc = time_e2e(zc_launch)
    torch.cuda.synchronize()
    # the two routes deliver the same bytes to the host (last timed step of each: same seed)
    assert torch.equal(h_zc_scores, h_scores) and torch.equal(h_zc_trace, h_trace) and torch.equal(h_zc_states, h_out_states)
    if e2e_zc_value >= e2e_copies_value:
        e2e_value, ms_e2e, e2e_route = e2e_zc_value, ms_e2e_zc, "zero-copy: one launch, the kernel reads / writes the pinned host buffers across PCIe itself"
    else:
        e2e_value, ms_e2e, e2e_route = e2e_copies_value, ms_e2e_copies, "cudaMemcpyAsync H2D + launch + three cudaMemcpyAsync D2H"

    # ---- per-launch kernels of the stepwise API + saturation sweep (explanatory, not the headline)
    extras = {}
    if rank == 0 and not args.quick and not args.no_sweep:
        extras = stepwise_and_sweep(torch, dev, BatchedGoEnv)

    # ---- MCTS half of the metric
    mcts = None
    if not args.no_mcts:
        sampler.active.set()
        mcts = mcts_bench(torch, dist, dev, rank, world, args)
        sampler.active.clear()
    c4 = None
    if rank == 0 and not args.quick and not args.no_sweep and not args.no_mcts:
        c4 = c4_bench(torch, dev, BatchedGoEnv)
    c5 = None
    if not args.no_mcts and not args.tree_only:
        c5 = c5_bench(torch, dist, dev, rank, world, args)
    sampler.stop()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    launch_ms = ms_kernel / K
    plies_per_launch = plies_all / world / K
    achieved = plies_per_launch * BYTES_PER_PLY / (launch_ms * 1e-3) / 1e9
    traffic = ncu_static = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_playout_traffic.json")))
        traffic, ncu_static = prof.get("dram_bytes_per_launch"), prof.get("ncu")
    except Exception:
        pass
    if rank == 0:
        line = {
            "metric": "board steps/sec (9x9 env-only batched random playouts) [+ MCTS sims/sec under 'mcts']",
            "value": value, "unit": "board steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 bitboards (board), f32/f64 (tree statistics)", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: 9x9 env-only random playouts, %d concurrent games per GPU, empty board "
                                   "to termination (max_step 120), hashed k-th-legal-point move rule, + area scoring" % G,
                       "games_per_gpu": G, "playout_layout": env.playout_layout() + (" (library choice)" if not env.tile else ""), "board": 9,
                       "l2": "256 MB flush buffer written between timed iterations", "sharding": "games by id, no collective"},
            "e2e": {"value": e2e_value, "unit": "board steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "route": e2e_route,
                    "routes": {"copies": {"value": e2e_copies_value, "ms_per_step": ms_e2e_copies / K},
                               "zero_copy": {"value": e2e_zc_value, "ms_per_step": ms_e2e_zc / K,
                                             "note": "one launch of tg_env_playout_io: pinned start positions read, final positions / scores / trace written by the kernel"}}},
            "gpu_launches": launches[0],
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": playout_kernel_name(env),
                         "launch_ms": launch_ms,
                         "bytes_per_ply": BYTES_PER_PLY, "plies_per_launch": plies_per_launch,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "ncu": ncu_static,
                         "note": "integer bit-manipulation kernel: state stays in registers for ~120 plies, so it is bound by the ALU "
                                 "pipe (IADD3/LOP3/SHF issue every other cycle per scheduler: 50 % issue slots is that pipe's ceiling) and "
                                 "the dependency chain of a game, not by HBM; 'ncu' = the static capture under profiles/"},
            "cpu_baseline": cpu,
        }
        line.update(extras)
        if mcts is not None:
            if cpu_mcts is not None:
                mcts["cpu_baseline"] = cpu_mcts
            if cpu_mcts_gpu_net is not None:
                mcts["reference_style_gpu_net"] = cpu_mcts_gpu_net
            line["mcts"] = mcts
        if c4 is not None:
            line["c4_19x19"] = c4
        if c5 is not None:
            line["c5"] = c5
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _bytes_per_sim(d_bar, k_bar):
    return d_bar * 24 * k_bar + 64 * (d_bar + 1) + 40 * k_bar + (4 * 82 + 4) + 16 + 2 * 48 + 3240


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def playout_kernel_name(env):
    t = env.lib.tg_env_playout_tile(env.n, env.G, env.tile)
    if t < 0:
        return "tg::k_playout_rows<%d>" % env.n
    if t & 128:
        return "tg::k_playout_inc<%d,%d,false>" % (env.n, t & 63)
    return ("tg::k_playout_wu<%d,%d>" if t & 64 else "tg::k_playout<%d,%d>") % (env.n, t & 63)


def stepwise_and_sweep(torch, dev, BatchedGoEnv):
    """Per-launch numbers of the stepwise API at G = 4096 and the playout saturation sweep."""
    out = {}

    def best_ms(fn, reps=5):
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    def first_legal(e):
        """a legal board move per game (the lowest legal point; PASS where none), computed outside the timed region"""
        m = e.unpack_legal(e.legal_mask())
        return m.to(torch.uint8).argmax(1).to(torch.int16)

    def step_ms(e, reps=5):
        best = 1e30
        for _ in range(reps):
            acts = first_legal(e)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); e.step(acts); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    env = BatchedGoEnv(GAMES, device=dev)
    env.playout(40, seed=2)
    ms = step_ms(env)
    out["step_kernel"] = {"games": GAMES, "ms": ms, "steps_per_s": GAMES / ms * 1e3, "gbps_algorithmic": GAMES * 115 / ms / 1e6}
    # a ply-by-ply driver: step + encode per ply, eager C-ABI calls vs one CUDA-graph replay per ply (host wall clock per ply)
    planes = torch.zeros([GAMES, 10, 9, 9], dtype=torch.float32, device=dev)
    env2 = BatchedGoEnv(GAMES, device=dev)
    acts0 = torch.full([GAMES], 40, dtype=torch.int16, device=dev)
    step_graph = env2.make_step_graph(planes=planes)
    def loop(fn, n=200):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e6
    def eager():
        env2.step(acts0); env2.encode(out=planes)
    loop(eager, 20); loop(lambda: step_graph(acts0), 20)
    out["stepwise_driver"] = {"games": GAMES, "what": "step + encode per ply (the reference's per-ply usage), host wall clock per ply",
                              "eager_us_per_ply": loop(eager), "graph_us_per_ply": loop(lambda: step_graph(acts0))}
    ms = best_ms(lambda: env.encode())
    out["encode_kernel"] = {"games": GAMES, "ms": ms, "positions_per_s": GAMES / ms * 1e3, "gbps_algorithmic": GAMES * 3288 / ms / 1e6}
    sweep = []
    for G, tile in ((4096, 0), (65536, 0), (1 << 20, 0)):
        e = BatchedGoEnv(G, device=dev, tile=tile)
        cnt = torch.zeros([1], dtype=torch.int64, device=dev)

        def play():
            e.reset(); cnt.zero_(); e.playout(e.max_step + 1, seed=1, counter=cnt)
        ms = best_ms(play, 3)
        row = {"games": G, "layout": e.playout_layout(), "ms": ms, "steps_per_s": int(cnt.item()) / ms * 1e3}
        if G >= 65536:                                         # the HBM-heavy kernels at saturation (mid-game positions)
            e.reset(); e.playout(60, seed=4)
            ms_s = step_ms(e, 3)
            buf = torch.empty([G, 10, 9, 9], dtype=torch.float32, device=dev)
            ms_e = best_ms(lambda: e.encode(out=buf), 3)
            row.update(step_kernel_per_s=G / ms_s * 1e3, step_kernel_gbps=G * 115 / ms_s / 1e6,
                       encode_per_s=G / ms_e * 1e3, encode_gbps=G * 3288 / ms_e / 1e6)
            del buf
        sweep.append(row)
        del e
    out["saturation_sweep"] = sweep
    return out


def c4_bench(torch, dev, BatchedGoEnv):
    """BASELINE configs[3] (board-size generalisation): 19x19, 512 games; env playouts and one tree-only search move of 400
    simulations.  Explanatory numbers, not the headline."""
    from transgo_b200 import _lib
    from transgo_b200.mcts import BatchedMCTS
    L = _lib.lib()
    G, sims = 512, 400
    env = BatchedGoEnv(G, board_size=19, device=dev)
    cnt = torch.zeros([1], dtype=torch.int64, device=dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env.playout(50, seed=3)
    env.reset(); torch.cuda.synchronize()
    a.record(); env.playout(env.max_step + 1, seed=1, counter=cnt); b.record(); torch.cuda.synchronize()
    out = {"config": "19x19, %d games, max_step %d; MCTS %d simulations/move" % (G, env.max_step, sims),
           "playout_steps_per_s": int(cnt.item()) / a.elapsed_time(b) * 1e3, "playout_ms": a.elapsed_time(b)}
    del env
    wave = [0]
    pol = {G * 4: torch.empty([G * 4, 362], dtype=torch.float32, device=dev), G: torch.empty([G, 362], dtype=torch.float32, device=dev)}
    val = {G * 4: torch.empty([G * 4], dtype=torch.float32, device=dev), G: torch.empty([G], dtype=torch.float32, device=dev)}

    def synth(obs):
        B = obs.shape[0]
        _lib.check(L.tg_synth_provider(19, B, 7, wave[0], _lib.ptr(pol[B]), _lib.ptr(val[B]), _lib.stream_ptr()))
        wave[0] += 1
        return pol[B], val[B]
    eng = BatchedMCTS(synth, G, board_size=19, num_simulation=sims, tie_seed=1, device=dev)
    actions, _, _ = eng.get_action_probs(True)
    eng.update_with_action(actions)
    eng.total_sims()
    torch.cuda.synchronize()
    a.record()
    actions, _, _ = eng.get_action_probs(True)
    eng.update_with_action(actions)
    b.record(); torch.cuda.synchronize()
    n = eng.total_sims()
    out["mcts_tree_only_sims_per_s"] = n / a.elapsed_time(b) * 1e3
    out["mcts_ms_per_move"] = a.elapsed_time(b)
    out["workspace_gb"] = eng.workspace_bytes / 1e9
    del eng
    # the same search with the random-init 19x19 TransGo net (fused bf16 forward; its attention blocks use the library
    # batched-GEMM path above 96 points), one move
    from transgo_b200.fused_net import make_fused_evaluator
    from transgo_b200.net import TransGoNetwork
    torch.manual_seed(0)
    net = TransGoNetwork(board_size=19).to(dev).eval()
    eng = BatchedMCTS(make_fused_evaluator(net, cuda_graph=True), G, board_size=19, num_simulation=sims, tie_seed=1, device=dev)
    for _ in range(2):
        eng.run_wave()
    eng.total_sims()
    torch.cuda.synchronize()
    a.record()
    actions, _, _ = eng.get_action_probs(True)
    eng.update_with_action(actions)
    b.record(); torch.cuda.synchronize()
    out["mcts_e2e_net_bf16_fused_sims_per_s"] = eng.total_sims() / a.elapsed_time(b) * 1e3
    out["mcts_e2e_ms_per_move"] = a.elapsed_time(b)
    return out


def c5_bench(torch, dist, dev, rank, world, args):
    """BASELINE configs[4]: the full self-play round on every rank of the job -- NCCL weight broadcast from the trainer rank
    (replaces the per-game weight fetch, self_play.py:913 / trainer.py:75-79), this rank's shard of games searched with the
    fused TransGo evaluator (1024 games x 800 simulations per move, a bounded number of moves), labelled positions gathered
    to the replay owner with one variable-size message per rank (replaces the 8.T appends per game, self_play.py:956-965),
    x8 augmentation and ring append on the owner.  Phase times are device times, max over ranks."""
    from transgo_b200.dist import shard_games
    from transgo_b200.fused_net import make_fused_evaluator
    from transgo_b200.net import TransGoNetwork
    from transgo_b200.replay import Pacing, ReplayRing
    from transgo_b200.selfplay import SelfPlayWorker, selfplay_round
    G, sims, moves = (MCTS_GAMES, MCTS_SIMS, 3) if not args.quick else (256, 64, 3)
    torch.manual_seed(1234 + rank)                             # ranks start with DIFFERENT weights: the broadcast must fix that
    model = TransGoNetwork(board_size=9).to(dev).eval()
    ev = make_fused_evaluator(model, cuda_graph="roots")
    game_id0, _ = shard_games(G * world, rank, world)
    worker = SelfPlayWorker(ev, G, num_simulation=sims, tie_seed=7, game_id0=game_id0, device=dev, graph_waves=True)
    ring = ReplayRing(1 << 18, device=dev) if rank == 0 else None
    pacing = Pacing() if rank == 0 else None
    out = None
    for rnd in range(2):                                       # round 0 warms up (graph capture, NCCL channels); round 1 is reported
        if rank == 0:
            with torch.no_grad():                              # stand-in for train steps: the trainer rank's weights move
                for p in model.parameters():
                    p.mul_(1.0 + 1e-3 * (rnd + 1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        worker.search.total_sims()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        st = selfplay_round(worker, model, ring, pacing, src=0, dst=0, max_moves=moves, refresh=ev.refresh, timing=True)
        b.record()
        torch.cuda.synchronize()
        n_sims = worker.search.total_sims()
        t = torch.tensor([a.elapsed_time(b), st["broadcast_ms"], st["play_ms"], st["gather_ms"], st["append_ms"]], dtype=torch.float64, device=dev)
        c = torch.tensor([n_sims, st["positions"], st["gather_bytes_sent"]], dtype=torch.float64, device=dev)
        digest = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        lo, hi = digest.clone(), digest.clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out = {"config": {"workload": "BASELINE configs[4]: 9x9 self-play round sharded by game batch, %d games x %d simulations/move per GPU, "
                                      "%d moves per round, fused bf16 TransGo evaluator, NCCL weight broadcast + trajectory gather" % (G, sims, moves),
                          "games_per_gpu": G, "num_simulation": sims, "moves": moves},
               "round_ms": float(t[0]), "broadcast_ms": float(t[1]), "play_ms": float(t[2]), "gather_ms": float(t[3]), "append_x8_ms": float(t[4]),
               "broadcast_bytes": st["broadcast_bytes"], "gather_bytes": float(c[2]), "gather_frac_of_round": float(t[3]) / float(t[0]),
               "positions": float(c[1]), "positions_per_s": float(c[1]) / (float(t[0]) * 1e-3), "mcts_sims_per_s": float(c[0]) / (float(t[0]) * 1e-3),
               "weights_identical_on_all_ranks": bool(float(lo[0]) == float(hi[0])),
               "rows_appended_x8": st["appended"] if rank == 0 else None,
               "collectives": "dist.broadcast (one flat buffer per dtype), all_gather_into_tensor (row counts), batch_isend_irecv (rows)" if world > 1 else "none (1 GPU)"}
    del worker, ring
    return out


def mcts_bench(torch, dist, dev, rank, world, args):
    """BASELINE configs[2]: 1024 games x 800 simulations per move.  (i) tree kernels alone: policy/value come from the
    device-side synthetic provider; (ii) end to end with the random-init TransGo net (PyTorch: bf16 weights + channels_last,
    and fp32), forward replayed as a CUDA graph on the engine's static input tensors."""
    from transgo_b200 import _lib
    from transgo_b200.mcts import BatchedMCTS
    from transgo_b200.net import TransGoNetwork, make_evaluator
    L = _lib.lib()
    G, sims = (MCTS_GAMES, MCTS_SIMS) if not args.quick else (256, 64)
    moves = 3 if not args.tree_only else 1
    out = {"config": {"workload": "BASELINE configs[2]: 9x9 batched MCTS self-play, %d games x %d simulations/move, 4 leaves per "
                                  "wave and game (configure.py:33), c_puct 3/0.05, wu_loss 2, Dirichlet(0.03) root noise" % (G, sims),
                      "games_per_gpu": G, "num_simulation": sims}}

    def timed_moves(eng, moves):
        eng.total_sims()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a.record()
        w0 = eng.waves
        for _ in range(moves):
            actions, pi, obs = eng.get_action_probs(True)
            eng.update_with_action(actions)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        n = eng.total_sims()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        s = torch.tensor([n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
        return float(s[0]) / (float(t[0]) * 1e-3), float(t[0]), float(s[0]), eng.waves - w0

    wave = [0]
    pol = torch.empty([G * 4, 82], dtype=torch.float32, device=dev)
    val = torch.empty([G * 4], dtype=torch.float32, device=dev)
    pol_r = torch.empty([G, 82], dtype=torch.float32, device=dev)
    val_r = torch.empty([G], dtype=torch.float32, device=dev)

    wave_dev = torch.zeros([2], dtype=torch.int64, device=dev)

    def synth(obs):                                            # capturable: the wave index lives on the device
        B = obs.shape[0]
        p, v = (pol, val) if B == G * 4 else (pol_r, val_r)
        _lib.check(L.tg_synth_provider_dev(9, B, 99, _lib.ptr(wave_dev), _lib.ptr(p), _lib.ptr(v), _lib.stream_ptr()))
        return p, v
    eng = BatchedMCTS(synth, G, num_simulation=sims, tie_seed=1, game_id0=rank * G, device=dev, graph_waves=True)
    timed_moves(eng, 1)                                        # warm-up move
    eng.search_stats()
    rate, ms, n, waves = timed_moves(eng, moves)
    st = eng.search_stats()
    d_bar, k_bar = st["mean_depth"], st["mean_fanout"]
    # SURVEY 8d traffic model with this engine's record sizes: 24 B per child scanned (16 B hot record + 8 B child info), 64 B
    # RMW per path node (hot + cold, read and written), 40 B per new child slot, policy row + value, legal mask, two states,
    # the 3240 B of planes
    bps = _bytes_per_sim(d_bar, k_bar)
    out["tree_only"] = {"value": rate, "unit": "MCTS sims/s", "ms": ms, "sims": n, "moves": moves, "waves": waves,
                        "provider": "tg_synth_provider (device-side synthetic policy/value, NOT the net); each wave (select, leaves, "
                                    "provider, expand) replayed as one CUDA graph",
                        "mean_depth": d_bar, "mean_fanout": k_bar, "bytes_per_sim": bps,
                        "gbps_algorithmic": rate * bps / 1e9 / world,
                        "roofline": {"bound": "hbm", "achieved": rate * bps / 1e9 / world, "peak": _hbm_peak(), "unit": "GB/s",
                                     "frac": rate * bps / 1e9 / world / _hbm_peak(),
                                     "note": "per GPU; algorithmic bytes per simulation from the measured depth / fan-out (DESIGN.md 4); "
                                             "the tree kernels are latency bound (one warp per game, dependent loads + float64 score chain), "
                                             "see profiles/r02_mcts_ncu_summary.txt"}}
    del eng
    # the same workload with the batch cut into shards whose waves overlap on separate CUDA streams (mcts.ConcurrentMCTS):
    # every tree kernel is latency bound at 1024 games, so independent shards fill the SMs the single batch leaves idle
    from transgo_b200.mcts import ConcurrentMCTS
    best = None
    for shards in ((2, 4, 8) if not args.quick else (4,)):
        per = G // shards
        bufs = [(torch.empty([per * 4, 82], dtype=torch.float32, device=dev), torch.empty([per * 4], dtype=torch.float32, device=dev),
                 torch.empty([per, 82], dtype=torch.float32, device=dev), torch.empty([per], dtype=torch.float32, device=dev),
                 torch.zeros([2], dtype=torch.int64, device=dev)) for _ in range(shards)]

        def make_synth(i):
            pw, vw, pr, vr, wv = bufs[i]

            def synth_i(obs):
                B = obs.shape[0]
                p, v = (pw, vw) if B == per * 4 else (pr, vr)
                _lib.check(L.tg_synth_provider_dev(9, B, 99 + i, _lib.ptr(wv), _lib.ptr(p), _lib.ptr(v), _lib.stream_ptr()))
                return p, v
            return synth_i
        ceng = ConcurrentMCTS([make_synth(i) for i in range(shards)], G, shards=shards, num_simulation=sims, tie_seed=1, game_id0=rank * G,
                              device=dev, graph_waves=True, waves_per_graph=4)
        timed_moves(ceng, 1)
        ceng.search_stats()
        rate_c, ms_c, n_c, waves_c = timed_moves(ceng, moves)
        st_c = ceng.search_stats()
        row = {"value": rate_c, "unit": "MCTS sims/s", "shards": shards, "ms": ms_c, "sims": n_c, "moves": moves, "waves": waves_c,
               "mean_depth": st_c["mean_depth"], "mean_fanout": st_c["mean_fanout"]}
        if best is None or rate_c > best["value"]:
            best = row
        out.setdefault("tree_only_streams_sweep", []).append({"shards": shards, "value": rate_c, "ms": ms_c})
        del ceng, bufs
    bps_c = _bytes_per_sim(best["mean_depth"], best["mean_fanout"])
    best.update(bytes_per_sim=bps_c, provider="tg_synth_provider per shard; the shards' waves (one CUDA graph each) are replayed round-robin on "
                                               "separate streams", roofline={"bound": "hbm", "achieved": best["value"] * bps_c / 1e9 / world, "peak": _hbm_peak(),
                                                                           "unit": "GB/s", "frac": best["value"] * bps_c / 1e9 / world / _hbm_peak()})
    out["tree_only_streams"] = best
    if args.tree_only:
        return out
    torch.manual_seed(0)
    net = TransGoNetwork(board_size=9).to(dev).eval()
    from transgo_b200.fused_net import make_fused_evaluator
    if rank == 0 and not args.quick:
        out["c1_single_game_facade"] = c1_facade(torch, net)
    for name, dtype in (("e2e_net_bf16_fused", "fused"), ("e2e_net_bf16", torch.bfloat16), ("e2e_net_fp32_folded", "folded32"),
                        ("e2e_net_fp32", None)):
        if args.quick and dtype in (None, "folded32"):
            continue
        # fused: the whole wave (tree kernels + net) is one CUDA graph, the root batch has its own; eager nets: graph per forward.
        # folded32: the fused SCHEDULE (BatchNorm folded, NHWC, one graph per wave) with float32 storage and library kernels -- the
        # same precision as the reference's net (float32, TF32 convolutions as PyTorch runs them by default)
        graphed = dtype in ("fused", "folded32")
        ev = make_fused_evaluator(net, cuda_graph="roots") if dtype == "fused" else \
            make_fused_evaluator(net, dtype=torch.float32, cuda_graph="roots") if dtype == "folded32" else \
            make_evaluator(net, dtype=dtype, channels_last=dtype is not None, cuda_graph=True)
        eng = BatchedMCTS(ev, G, num_simulation=sims, tie_seed=1, game_id0=rank * G, device=dev, graph_waves=graphed)
        # warm-up waves: lazy initialisation and CUDA-graph capture happen here, outside the timed moves (they only add visits to the
        # fresh roots); the graphed routes warm up for ~0.1-0.3 s so that the timed region does not start on clocks that are still
        # ramping up after the host-bound phases before it (a 0.55 s measurement read 20 % low when it did)
        for _ in range(60 if graphed else 3):
            eng.run_wave()
        nmoves = 2 if dtype == "fused" else 1
        rate, ms, n, waves = timed_moves(eng, nmoves)
        out[name] = {"value": rate, "unit": "MCTS sims/s", "ms": ms, "sims": n, "moves": nmoves, "waves": waves,
                     "net": "TransGoNetwork 2.97 M params random init (seed 0), eval mode, 0.457 GFLOP/position" +
                            ("; BN folded, NHWC bf16; hand-written: stem (tg_net_stem), residual junctions (tg_net_add_bn_relu), attention "
                             "blocks incl. their 1x1 convolutions (tg_net_attention_fused2: tcgen05), head convolutions "
                             "(tg_net_conv3x3_head); the 18 trunk convolutions in cuDNN, the head GEMMs in cuBLAS" if dtype == "fused" else
                             "; BN folded, NHWC float32, library kernels only (float32 storage, TF32 convolutions)" if dtype == "folded32" else ""),
                     "net_tflops": rate * 0.457e9 / 1e12}
        del eng
    return out


def c1_facade(torch, net):
    """BASELINE configs[0] through the drop-in facade: ONE game, the reference's default search (configure.py:25-33: 210
    simulations, 4 leaves per wave), the scalar WP_MCTS / GoEnv surface of self_play.py:906-926, fp32 net in eager mode --
    what a user gets by only swapping the two imports.  Three moves after one warm-up move."""
    from transgo_b200.environment import GoEnv
    from transgo_b200.mcts import WP_MCTS

    class Cfg:
        board_size, num_simulation, c_puct1, c_puct2, wu_loss, parallel_readouts = 9, 210, 3, 0.05, 2, 4
    agent = WP_MCTS(Cfg(), GoEnv(Cfg()), net, None)
    a, _, _ = agent.get_action_probs(True)
    agent.update_with_action(a)
    agent._b.total_sims()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        a, _, _ = agent.get_action_probs(True)
        agent.update_with_action(a)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = agent._b.total_sims()
    return {"value": n / dt, "unit": "MCTS sims/s", "moves": 3, "sims": n, "wall_s": dt,
            "what": "one game, 210 simulations per move, WP_MCTS facade + fp32 eager net on the GPU (host wall clock)"}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 1)
    cores = os.cpu_count() or 1
    per = 0.25                                                 # seconds per step and core: bounded sample
    for _ in range(min(W, 2)):
        cpu_playout_baseline(per, cores)
    plies, wall, kind, games = 0.0, 0.0, "reference", 0
    for _ in range(K):
        r, (n, ng, w) = cpu_playout_baseline(per, cores)
        kind = r["kind"]
        plies += n
        wall += w
        games += ng
    value = plies / wall
    line = {"impl": "reference", "metric": "board steps/sec (9x9 env-only batched random playouts) [+ MCTS sims/sec under 'mcts']",
            "value": value, "unit": "board steps/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": wall / K * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int (reference C++ engine)", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: 9x9 env-only random playouts, same hashed move rule, whole games; each step is "
                                   "a bounded sample of %.2f s on every host core" % per, "board": 9},
            "cpu_baseline": {"value": value, "unit": "board steps/s", "cores": cores, "kind": kind,
                             "sample": "%d steps x %d processes x %.2f s: %d games, %d plies" % (K, cores, per, games, int(plies))},
            "e2e": {"value": value, "unit": "board steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_mcts:
        # second half of the metric on the host: BASELINE configs[0] defaults, one game per process on every core
        m = cpu_mcts_baseline(cores=cores, sims=210)
        line["mcts"] = {"config": {"workload": "BASELINE configs[0]: 9x9 self-play search, reference defaults (210 simulations/move, 4 leaves "
                                               "per net call), random-init TransGo net on the CPU, one game per process x %d processes" % cores},
                        "value": m["value"], "unit": "MCTS sims/s", "cpu_baseline": m}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=GAMES)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--no-mcts", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="small MCTS phase, no sweep (profiling runs)")
    ap.add_argument("--tree-only", action="store_true", help="MCTS phase: skip the end-to-end net runs (profiling runs)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the stepwise kernels / saturation sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
