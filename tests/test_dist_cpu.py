"""CPU tests (gloo, world_size 2) of the off-hot-path plumbing of SURVEY 8e, and of the trajectory augmentation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transgo_b200.dist import broadcast_weights, gather_trajectories, shard_games
    from transgo_b200.net import TransGoNetwork
    torch.manual_seed(100 + rank)                     # different weights per rank before the broadcast
    net = TransGoNetwork(board_size=9, num_features=16)
    broadcast_weights(net, src=0)
    digest = float(sum(v.double().sum() for v in net.state_dict().values()))
    g0, cnt = shard_games(1025, rank, world)
    n = 3 + 2 * rank                                  # ragged: ranks finish with different position counts
    obs = torch.full([n, 10, 9, 9], float(rank)); pi = torch.full([n, 82], 0.5 + rank)
    z = torch.full([n], 1.0 - 2 * rank); own = torch.full([n, 81], float(-rank))
    got = gather_trajectories(obs, pi, z, own, dst=0)
    empty = gather_trajectories(obs[:0], pi[:0], z[:0], own[:0], dst=0) if True else None
    # uneven shards: the row-count exchange also carries each rank's game count (selfplay_round sums them for the pacing counters)
    got2, extras = gather_trajectories(obs, pi, z, own, dst=1, extra=[cnt, 7 + rank])
    res = dict(rank=rank, digest=digest, shard=(g0, cnt), extras=extras.tolist(), got2=None if got2 is None else tuple(got2[0].shape))
    if rank == 0:
        o, p, zz, w = got
        res.update(shapes=(tuple(o.shape), tuple(p.shape), tuple(zz.shape), tuple(w.shape)),
                   ranks_seen=sorted(set(o[:, 0, 0, 0].tolist())), z=zz.tolist(), empty=tuple(empty[0].shape))
    q.put(res)
    dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r["rank"])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0]["digest"] == res[1]["digest"]                      # rank 1 now holds rank 0's weights
    assert res[0]["shard"] == (0, 513) and res[1]["shard"] == (513, 512)
    assert res[0]["shapes"] == ((8, 10, 9, 9), (8, 82), (8,), (8, 81))
    assert res[0]["ranks_seen"] == [0.0, 1.0] and res[0]["z"] == [1.0] * 3 + [-1.0] * 5
    assert res[0]["empty"] == (0, 10, 9, 9)
    assert res[0]["extras"] == res[1]["extras"] == [[513, 7], [512, 8]]     # every rank sees every rank's counts
    assert res[0]["got2"] is None and res[1]["got2"] == (8, 10, 9, 9)        # any rank can be the replay owner


def test_augment8_matches_the_reference_loop():
    """self_play.py:943-965 restated with numpy (rot90 by i then fliplr of that rotation), compared to the torch batch op."""
    from transgo_b200.selfplay import augment8
    rng = np.random.default_rng(0)
    B, n = 3, 9
    obs = rng.random((B, 10, n, n)).astype(np.float32); pi = rng.random((B, n * n + 1)).astype(np.float32)
    z = np.array([1, -1, 1], np.float32); own = rng.integers(-1, 2, (B, n * n)).astype(np.float32)
    exp_o, exp_p, exp_z, exp_w = [], [], [], []
    for o, p, zz, w in zip(obs, pi, z, own):
        for i in (1, 2, 3, 4):
            ep = np.rot90(p[:-1].reshape(n, n), i)
            eo = np.array([np.rot90(s, i) for s in o]); ew = np.rot90(w.reshape(n, n), i)
            exp_o.append(eo); exp_p.append(np.append(ep.flatten(), p[-1])); exp_z.append(zz); exp_w.append(ew.flatten())
            exp_o.append(np.array([np.fliplr(s) for s in eo])); exp_p.append(np.append(np.fliplr(ep).flatten(), p[-1]))
            exp_z.append(zz); exp_w.append(np.fliplr(ew).flatten())
    o, p, zz, w = augment8(torch.from_numpy(obs), torch.from_numpy(pi), torch.from_numpy(z), torch.from_numpy(own), n)
    assert np.array_equal(o.numpy(), np.array(exp_o)) and np.array_equal(p.numpy(), np.array(exp_p))
    assert np.array_equal(zz.numpy(), np.array(exp_z)) and np.array_equal(w.numpy(), np.array(exp_w))


def test_net_is_state_dict_compatible_with_the_reference():
    """226 tensors / 2,977,512 elements; live comparison when /root/reference is present."""
    from transgo_b200.net import TransGoNetwork
    torch.manual_seed(0)
    mine = TransGoNetwork(board_size=9).eval()
    sd = mine.state_dict()
    assert len(sd) == 226 and sum(v.numel() for v in sd.values()) == 2977512
    if not os.path.exists("/root/reference/model.py"):
        return
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_model", "/root/reference/model.py")
    rm = importlib.util.module_from_spec(spec); spec.loader.exec_module(rm)

    class C:
        board_size, input_dim, num_features = 9, 10, 128
    ref = rm.TransGoNetwork(C()).eval()
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, .1); m.running_var.uniform_(.5, 1.5)
        for k, p in ref.named_parameters():
            if k.endswith("gamma"):
                p.fill_(0.3)
    mine.set_weights(ref.get_weights())
    x = (torch.rand(4, 10, 9, 9) > 0.7).float()
    with torch.no_grad():
        for a, b in zip(ref.main_prediction(x), mine.main_prediction(x)):
            assert torch.equal(a, b)


def test_replay_ring_and_pacing_cpu(tmp_path):
    """replay_buffer.py:16-94 semantics (ring index / full flag / sampling rule / pickle layout) and the
    shared_storage.py + configure.py schedules."""
    from transgo_b200.replay import ReplayRing, Pacing
    r = ReplayRing(10, device="cpu")
    mk = lambda b, v: (torch.full([b, 10, 9, 9], float(v)), torch.full([b, 82], float(v)), torch.full([b], float(v)), torch.full([b, 81], float(v)))
    r.append(*mk(4, 1)); assert r.info() == {"capacity": 10, "index": 4, "full": False} and len(r) == 4
    r.append(*mk(8, 2)); assert r.info() == {"capacity": 10, "index": 2, "full": True} and len(r) == 10
    assert r.z.tolist() == [2, 2, 1, 1, 2, 2, 2, 2, 2, 2]                 # wrapped rows overwrote the oldest
    o, p, z, w = r.sample(6)
    assert o.shape == (6, 10, 9, 9) and p.shape == (6, 82) and z.shape == (6,) and w.shape == (6, 81)
    small = ReplayRing(10, device="cpu"); small.append(*mk(3, 5))
    assert small.sample(7)[2].tolist() == [5.0] * 7                        # fewer rows than the batch: sample with replacement
    path = str(tmp_path / "replay_buffer0.pkl")
    saved = r.save(path)
    assert set(saved) == {"buffer_capacity", "index", "full", "data", "save_len"} and saved["data"].dtype == object
    assert saved["data"][0][0].shape == (10, 9, 9) and saved["data"][0][1].shape == (82,) and saved["data"][0][3].shape == (81,)
    r2 = ReplayRing(10, device="cpu"); r2.load(path)
    assert torch.equal(r2.z.sort().values, r.z.sort().values)
    pc = Pacing()
    pc.add_play_steps(11); assert pc.train_play_ratio == 7500 / 100000     # no change before the first finished game
    pc.add_play_games(1); pc.add_play_steps(1)
    assert pc.now_play_steps == 12 and pc.train_play_ratio == 7500 / 100000
    pc.add_play_steps(5)                                                   # step 17: (17 + 1) % 6 == 0 -> +1e-5
    assert abs(pc.train_play_ratio - (7500 + 1) / 100000) < 1e-12
    assert pc.selfplay_should_wait() and not pc.trainer_should_wait()
    pc.add_play_games(1498); assert pc.learn_rate == 6.5e-5 / 2            # game 1499: (1499 + 1) % 1500 == 0
