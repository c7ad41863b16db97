"""GPU-resident replay ring with the reference's ReplayMemory_Random semantics (replay_buffer.py:16-94) and counters
that stand in for the SharedStorage pacing fields (shared_storage.py:13-43).

The reference stores 1.5 M python tuples (obs[10,n,n], pi[A], z, own[P]) in an object array inside a Ray actor; here the
four fields are preallocated device tensors written by whole batches (the output of SelfPlayWorker / gather_trajectories),
sampled with one index_select per field, and saved in the reference's pickle layout so either side can load the other's file.
"""
import pickle

import numpy as np
import torch


class ReplayRing:
    def __init__(self, capacity, board_size=9, channels=10, device="cuda", store_dtype=torch.float32):
        n = board_size
        self.capacity, self.n, self.channels = int(capacity), n, channels
        self.device = torch.device(device)
        self.obs = torch.zeros([self.capacity, channels, n, n], dtype=store_dtype, device=self.device)
        self.pi = torch.zeros([self.capacity, n * n + 1], dtype=torch.float32, device=self.device)
        self.z = torch.zeros([self.capacity], dtype=torch.float32, device=self.device)
        self.own = torch.zeros([self.capacity, n * n], dtype=torch.float32, device=self.device)
        self.index, self.full = 0, False                       # replay_buffer.py:21-23

    def __len__(self):
        return self.capacity if self.full else self.index

    def append(self, obs, pi, z, own):
        """Batch form of replay_buffer.py:30-34: rows are written at index, index+1, ... modulo capacity."""
        b = obs.shape[0]
        if b == 0:
            return
        if b > self.capacity:
            obs, pi, z, own = obs[-self.capacity:], pi[-self.capacity:], z[-self.capacity:], own[-self.capacity:]
            b = self.capacity
        pos = (torch.arange(b, device=self.device) + self.index) % self.capacity
        self.obs[pos] = obs.to(self.obs.dtype); self.pi[pos] = pi.float(); self.z[pos] = z.float(); self.own[pos] = own.float()
        self.full = self.full or (self.index + b >= self.capacity)
        self.index = (self.index + b) % self.capacity

    def append_games(self, hist_obs, hist_pi, hist_player, length, winner, terr, fin=None, augment=True):
        """Finished games -> labelled (x8) rows written by ONE kernel launch straight into the ring's tensors at the write
        position (csrc/tg_traj.cu; self_play.py:931-965 + replay_buffer.py:30-34).  Returns the number of rows appended."""
        from .selfplay import assemble_trajectories
        assert self.obs.dtype == torch.float32, "append_games writes float32 rows"
        *_, rows = assemble_trajectories(hist_obs, hist_pi, hist_player, length, winner, terr, fin=fin, augment=augment,
                                         out=(self.obs, self.pi, self.z, self.own), index=self.index)
        self.full = self.full or (self.index + rows >= self.capacity)
        self.index = (self.index + rows) % self.capacity
        return rows

    def sample(self, batch_size, generator=None):
        """replay_buffer.py:36-47: without replacement when enough data, with replacement otherwise."""
        n = len(self)
        if n == 0:
            raise ValueError("empty replay ring")
        if n < batch_size:
            idx = torch.randint(0, n, [batch_size], device=self.device, generator=generator)
        else:
            idx = torch.randperm(n, device=self.device, generator=generator)[:batch_size]
        return self.obs[idx].float(), self.pi[idx], self.z[idx], self.own[idx]

    def info(self):                                            # replay_buffer.py:89-94
        return {"capacity": self.capacity, "index": self.index, "full": self.full}

    # ---- reference pickle layout (replay_buffer.py:49-87, transgo.py:225-250)
    def save(self, path=None):
        """The dict ReplayMemory_Random.save() returns for capacity <= 1.5 M (replay_buffer.py:51-56).  `data` is what
        `np.array([blank_trans] * capacity, dtype=object)` (replay_buffer.py:27) makes of a list of 4-tuples: an object array
        of shape [capacity, 4] whose row i holds (obs float64 [C,n,n], pi float64 [A], z float, own float64 [P]); unwritten
        rows keep the reference's blank transition.  Built column by column from whole-array host copies (no per-row device
        traffic, no Python-level row loop)."""
        n = len(self)
        o, p, z, w = (t[:n].detach().cpu().numpy() for t in (self.obs.float(), self.pi, self.z, self.own))
        o, p, w = o.astype(np.float64), p.astype(np.float64), w.astype(np.float64)
        data = np.empty((self.capacity, 4), dtype=object)
        if n:
            data[:n, 0] = np.fromiter(iter(o), dtype=object, count=n)      # row views of one host array
            data[:n, 1] = np.fromiter(iter(p), dtype=object, count=n)
            data[:n, 2] = z.astype(np.float64).tolist()
            data[:n, 3] = np.fromiter(iter(w), dtype=object, count=n)
        if n < self.capacity:
            blank = (np.zeros((self.channels, self.n, self.n)), np.zeros(self.n * self.n + 1), 0.0, np.zeros(self.n * self.n))
            for c in range(4):
                col = np.empty(self.capacity - n, dtype=object)
                col.fill(blank[c])
                data[n:, c] = col
        out = {"buffer_capacity": self.capacity, "index": self.index, "full": self.full, "data": data, "save_len": self.capacity}
        if path:
            with open(path, "wb") as f:
                pickle.dump(out, f)
        return out

    def load(self, info):
        """replay_buffer.py:72-87 for either writer's file.  A whole-capacity file (the <= 1.5 M branch, ours or the
        reference's) restores index / full exactly; a partial file (the reference's > 1.5 M branch: the last save_len rows)
        is appended at the current write position like the reference's load_index."""
        if isinstance(info, str):
            with open(info, "rb") as f:
                info = pickle.load(f)
        save_len = int(info["save_len"])
        rows = info["data"][:save_len]
        whole = "full" in info and save_len == int(info["buffer_capacity"])
        if whole:
            n_valid = save_len if info["full"] else int(info["index"])
            rows = rows[:n_valid]
        if len(rows) == 0:
            return self.full
        obs = torch.from_numpy(np.stack([r[0] for r in rows]).astype(np.float32)).to(self.device)
        pi = torch.from_numpy(np.stack([r[1] for r in rows]).astype(np.float32)).to(self.device)
        z = torch.tensor([float(r[2]) for r in rows], dtype=torch.float32, device=self.device)
        own = torch.from_numpy(np.stack([r[3] for r in rows]).astype(np.float32)).to(self.device)
        if whole and save_len == self.capacity:
            k = obs.shape[0]
            self.obs[:k] = obs.to(self.obs.dtype); self.pi[:k] = pi; self.z[:k] = z; self.own[:k] = own
            self.index, self.full = int(info["index"]), bool(info["full"])
        else:
            self.append(obs, pi, z, own)
        return self.full


class Pacing:
    """The counters and schedules SharedStorage + Config drive (shared_storage.py:21-43, configure.py:82-103)."""

    def __init__(self, train_play_ratio=7500 / 100000, learn_rate=6.5e-5, adjust_lr=True, adjust_ratio=True):
        self.now_play_steps = self.now_play_games = self.now_train_steps = 0
        self.train_play_ratio, self.learn_rate = train_play_ratio, learn_rate
        self.adjust_lr, self.adjust_ratio = adjust_lr, adjust_ratio

    def add_play_steps(self, n=1):
        for _ in range(int(n)):
            self.now_play_steps += 1
            if self.adjust_ratio and self.now_play_games > 0 and (self.now_play_steps + 1) % 6 == 0 \
                    and self.train_play_ratio < 2.6 / 10:          # shared_storage.py:36-42, configure.py:95-103
                self.train_play_ratio = (self.train_play_ratio * 100000 + 1) / 100000

    def add_play_games(self, n=1):
        for _ in range(int(n)):
            self.now_play_games += 1
            g = self.now_play_games
            if self.adjust_lr and (g + 1) % 1500 == 0 and g < 3100 and self.learn_rate > 0.5 * 0.5 * 6.5e-5:     # configure.py:88-91
                self.learn_rate *= 0.5

    def selfplay_should_wait(self, game_total_num=None):       # self_play.py:970-980
        """True while self-play is ahead of the trainer: train/play ratio below target, ratio adjustment on, and (when
        given) fewer games played than game_total_num -- the three conjuncts of the reference's wait loop."""
        if not self.adjust_ratio:
            return False
        if game_total_num is not None and self.now_play_games >= game_total_num:
            return False
        return self.now_train_steps / max(1, self.now_play_steps) < self.train_play_ratio

    def finish_game(self, positions):
        """Counter updates of ONE finished game in the reference's order: now_play_steps once per position as it is played
        (self_play.py:921), then now_play_games (self_play.py:967)."""
        self.add_play_steps(positions)
        self.add_play_games(1)

    def state_dict(self):
        return dict(now_play_steps=self.now_play_steps, now_play_games=self.now_play_games, now_train_steps=self.now_train_steps,
                    train_play_ratio=self.train_play_ratio, learn_rate=self.learn_rate)

    def load_state_dict(self, d):
        for k, v in d.items():
            setattr(self, k, v)

    def trainer_should_wait(self):                             # trainer.py:94-105
        return self.now_train_steps / max(1, self.now_play_steps) > self.train_play_ratio


# ---- checkpoint pickle (transgo.py:171-184 writes it, transgo.py:70-76 reads it): {"weights", "evaluate_weights"}, each a
# CPU state_dict as model.get_weights() returns it (model.py:23-38)
def save_checkpoint(path, model, evaluate_model=None):
    """Writes current_policy.model / best_policy_*.model in the reference's layout; loadable by transgo.py unchanged."""
    weights = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ev = weights if evaluate_model is None else {k: v.detach().cpu() for k, v in evaluate_model.state_dict().items()}
    out = {"weights": weights, "evaluate_weights": ev}
    with open(path, "wb") as f:
        pickle.dump(out, f)
    return out


def load_checkpoint(path, model=None, evaluate_model=None):
    """Reads a reference (or our) checkpoint pickle; optionally loads "weights" into model and "evaluate_weights" into
    evaluate_model via set_weights / load_state_dict.  Returns the dict."""
    with open(path, "rb") as f:
        ck = pickle.load(f)
    if not isinstance(ck, dict) or "weights" not in ck or "evaluate_weights" not in ck:
        raise ValueError("%s is not a Transgo checkpoint: expected a dict with 'weights' and 'evaluate_weights'" % path)
    if model is not None:
        model.load_state_dict(ck["weights"])
    if evaluate_model is not None:
        evaluate_model.load_state_dict(ck["evaluate_weights"])
    return ck
