"""CPU tests of the data formats either side of the hot path: the replay pickle and the checkpoint
pickle must be interchangeable with the reference's own classes, and the pacing counters must follow its schedules.
Where /root/reference is present the reference classes themselves are the checker (ray stubbed); elsewhere the layout
checks still run."""
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
have_ref = os.path.exists(os.path.join(REF, "replay_buffer.py"))


def _ref_modules():
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")

        def remote(*args, **kwargs):
            if len(args) == 1 and callable(args[0]) and not kwargs:
                return args[0]
            return lambda f: f
        ray.remote, ray.get, ray.init = remote, (lambda x: x), (lambda *a, **k: None)
        sys.modules["ray"] = ray
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import configure
    import replay_buffer
    import shared_storage
    return configure, replay_buffer, shared_storage


def _rows(n, seed):
    rng = np.random.default_rng(seed)
    obs = (rng.random((n, 10, 9, 9)) > 0.6).astype(np.float32)
    pi = rng.random((n, 82)).astype(np.float32); pi /= pi.sum(1, keepdims=True)
    z = rng.choice([-1.0, 1.0], n).astype(np.float32)
    own = rng.integers(-1, 2, (n, 81)).astype(np.float32)
    return obs, pi, z, own


def test_checkpoint_pickle_layout_round_trip(tmp_path):
    """transgo.py:171-184 writes {"weights", "evaluate_weights"} (CPU state_dicts), transgo.py:70-76 reads it."""
    from transgo_b200.net import TransGoNetwork
    from transgo_b200.replay import load_checkpoint, save_checkpoint
    torch.manual_seed(1)
    a, b = TransGoNetwork(board_size=9, num_features=16), TransGoNetwork(board_size=9, num_features=16)
    path = str(tmp_path / "current_policy.model")
    save_checkpoint(path, a, evaluate_model=b)
    with open(path, "rb") as f:
        raw = pickle.load(f)
    assert set(raw) == {"weights", "evaluate_weights"} and all(not v.is_cuda for v in raw["weights"].values())
    c, d = TransGoNetwork(board_size=9, num_features=16), TransGoNetwork(board_size=9, num_features=16)
    load_checkpoint(path, model=c, evaluate_model=d)
    for (k, x), (_, y) in zip(a.state_dict().items(), c.state_dict().items()):
        assert torch.equal(x, y), k
    for (k, x), (_, y) in zip(b.state_dict().items(), d.state_dict().items()):
        assert torch.equal(x, y), k
    with open(path, "wb") as f:
        pickle.dump({"weights": 1}, f)
    with pytest.raises(ValueError):
        load_checkpoint(path)


@pytest.mark.skipif(not have_ref, reason="/root/reference absent")
def test_checkpoint_interchanges_with_the_reference_model(tmp_path):
    """A checkpoint written here loads into the reference's TransGoNetwork.set_weights, and one written the reference's
    way (model.get_weights() under both keys) loads here: identical forward outputs."""
    import importlib.util
    from transgo_b200.net import TransGoNetwork
    from transgo_b200.replay import load_checkpoint, save_checkpoint
    spec = importlib.util.spec_from_file_location("ref_model_ck", os.path.join(REF, "model.py"))
    rm = importlib.util.module_from_spec(spec); spec.loader.exec_module(rm)

    class C:
        board_size, input_dim, num_features = 9, 10, 128
    torch.manual_seed(3)
    mine = TransGoNetwork(board_size=9).eval()
    path = str(tmp_path / "best_policy_1.model")
    save_checkpoint(path, mine)
    ref = rm.TransGoNetwork(C()).eval()
    with open(path, "rb") as f:
        ref.set_weights(pickle.load(f)["weights"])                     # transgo.py:74-76 + self_play.py:913
    x = (torch.rand(3, 10, 9, 9) > 0.7).float()
    with torch.no_grad():
        for p, q in zip(ref.main_prediction(x), mine.main_prediction(x)):
            assert torch.equal(p, q)
    torch.manual_seed(4)
    ref2 = rm.TransGoNetwork(C()).eval()
    path2 = str(tmp_path / "current_policy.model")
    with open(path2, "wb") as f:                                       # transgo.py:176-179
        pickle.dump({"weights": ref2.get_weights(), "evaluate_weights": ref2.get_weights()}, f)
    mine2 = TransGoNetwork(board_size=9).eval()
    load_checkpoint(path2, model=mine2)
    with torch.no_grad():
        for p, q in zip(ref2.main_prediction(x), mine2.main_prediction(x)):
            assert torch.equal(p, q)


@pytest.mark.skipif(not have_ref, reason="/root/reference absent")
def test_replay_files_interchange_with_the_reference_replay_memory(tmp_path):
    """replay_buffer.py:49-87: the reference's ReplayMemory_Random loads a file ReplayRing.save() wrote, and ReplayRing loads
    a file the reference wrote (same rows, same index / full)."""
    configure, replay_buffer, _ = _ref_modules()
    from transgo_b200.replay import ReplayRing

    class Cfg:
        buffer_size, encode_state_channels, board_size = 64, 10, 9
    obs, pi, z, own = _rows(37, 0)
    ring = ReplayRing(64, device="cpu")
    ring.append(*(torch.from_numpy(t) for t in (obs, pi, z, own)))
    path = str(tmp_path / "replay_buffer0.pkl")
    ring.save(path)
    mem = replay_buffer.ReplayMemory_Random(Cfg())
    with open(path, "rb") as f:
        mem.load(pickle.load(f))                                        # transgo.py:240-247
    for i in range(37):
        o, p, zz, w = mem.data[i]
        assert np.array_equal(o, obs[i]) and np.array_equal(p.astype(np.float32), pi[i]) and zz == z[i] and np.array_equal(w, own[i])
    o, p, zz, w = mem.data[50]                                          # unwritten slots hold the reference's blank transition
    assert o.shape == (10, 9, 9) and not o.any() and p.shape == (82,) and zz == 0.0 and w.shape == (81,)
    batch = mem.sample(8)                                               # and the reference can sample / unpack them (trainer.py)
    assert len(batch) == 8 and batch[0][0].shape == (10, 9, 9)
    # the other direction: rows appended by the reference, saved by the reference
    mem2 = replay_buffer.ReplayMemory_Random(Cfg())
    obs2, pi2, z2, own2 = _rows(70, 1)                                  # 70 > 64: the ring wraps, full = True, index = 6
    for i in range(70):
        mem2.append(obs2[i].astype(np.float64), pi2[i].astype(np.float64), float(z2[i]), own2[i].astype(np.float64))
    saved = mem2.save()
    path2 = str(tmp_path / "replay_buffer1.pkl")
    with open(path2, "wb") as f:
        pickle.dump(saved, f)
    ring2 = ReplayRing(64, device="cpu")
    assert ring2.load(path2) is True
    assert ring2.info() == {"capacity": 64, "index": mem2.index, "full": mem2.full} == {"capacity": 64, "index": 6, "full": True}
    for i in range(64):
        src = i + 64 if i < 6 else i                                    # slots 0..5 were overwritten by rows 64..69
        assert np.array_equal(ring2.obs[i].numpy(), obs2[src]) and np.array_equal(ring2.pi[i].numpy(), pi2[src])
        assert float(ring2.z[i]) == z2[src] and np.array_equal(ring2.own[i].numpy(), own2[src])
    # a partly filled reference file keeps its fill level
    mem3 = replay_buffer.ReplayMemory_Random(Cfg())
    for i in range(5):
        mem3.append(obs2[i].astype(np.float64), pi2[i].astype(np.float64), float(z2[i]), own2[i].astype(np.float64))
    ring3 = ReplayRing(64, device="cpu"); ring3.load(mem3.save())
    assert ring3.info() == {"capacity": 64, "index": 5, "full": False} and len(ring3) == 5


@pytest.mark.skipif(not have_ref, reason="/root/reference absent")
def test_pacing_follows_the_reference_shared_storage():
    """shared_storage.py:21-43 + configure.py:88-103 driven with the self-play actor's call order (self_play.py:921, :967):
    the same sequence of finished games must leave the same counters, learning rate and train/play ratio."""
    configure, _, shared_storage = _ref_modules()
    from transgo_b200.replay import Pacing
    cfg = configure.Config()
    ck = {"now_play_steps": 0, "now_play_games": 0, "now_train_steps": 0, "learn_rate": cfg.learn_rate if hasattr(cfg, "learn_rate") else 6.5e-5,
          "adjust_lr": True, "train_play_ratio": cfg.train_play_ratio, "adjust_train_play_ratio": True}
    ss = shared_storage.SharedStorage(ck, cfg)
    pc = Pacing(train_play_ratio=cfg.train_play_ratio, learn_rate=ck["learn_rate"])
    rng = np.random.default_rng(5)
    for game in range(3200):
        n = int(rng.integers(30, 121))
        for _ in range(n):
            ss.set_info("now_play_steps")
        ss.set_info("now_play_games")
        pc.finish_game(n)
        if game % 400 == 0 or game > 3190:
            assert pc.now_play_steps == ss.get_info("now_play_steps") and pc.now_play_games == ss.get_info("now_play_games")
            assert pc.learn_rate == ss.get_info("learn_rate"), game
            assert pc.train_play_ratio == ss.get_info("train_play_ratio"), game
    assert pc.learn_rate < ck["learn_rate"]                              # the halvings really happened
    # the throttle predicate of self_play.py:970-980
    want = ss.get_info("now_train_steps") / max(1, ss.get_info("now_play_steps")) < ss.get_info("train_play_ratio")
    assert pc.selfplay_should_wait() == want is True
    assert pc.selfplay_should_wait(game_total_num=100) is False          # ... and now_play_games < game_total_num
    pc.now_train_steps = int(pc.now_play_steps * pc.train_play_ratio) + 1
    assert pc.selfplay_should_wait() is False and pc.trainer_should_wait() is True
