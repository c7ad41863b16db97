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
        n = len(self)
        data = np.empty(self.capacity, dtype=object)
        o, p, z, w = (t.cpu().numpy() for t in (self.obs.float(), self.pi, self.z, self.own))
        blank = (np.zeros((self.channels, self.n, self.n)), np.zeros(self.n * self.n + 1), 0.0, np.zeros(self.n * self.n))
        for i in range(self.capacity):
            data[i] = (o[i], p[i], float(z[i]), w[i]) if i < n or self.full else blank
        out = {"buffer_capacity": self.capacity, "index": self.index, "full": self.full, "data": data, "save_len": self.capacity}
        if path:
            with open(path, "wb") as f:
                pickle.dump(out, f)
        return out

    def load(self, info):
        if isinstance(info, str):
            with open(info, "rb") as f:
                info = pickle.load(f)
        rows = info["data"][:info["save_len"]]
        n_valid = info["save_len"] if info.get("full", False) else info.get("index", info["save_len"])
        rows = rows[:n_valid] if n_valid else rows
        if len(rows) == 0:
            return self.full
        obs = torch.from_numpy(np.stack([r[0] for r in rows]).astype(np.float32)).to(self.device)
        pi = torch.from_numpy(np.stack([r[1] for r in rows]).astype(np.float32)).to(self.device)
        z = torch.tensor([float(r[2]) for r in rows], dtype=torch.float32, device=self.device)
        own = torch.from_numpy(np.stack([r[3] for r in rows]).astype(np.float32)).to(self.device)
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

    def selfplay_should_wait(self):                            # self_play.py:970-980
        return self.now_train_steps / max(1, self.now_play_steps) < self.train_play_ratio

    def trainer_should_wait(self):                             # trainer.py:94-105
        return self.now_train_steps / max(1, self.now_play_steps) > self.train_play_ratio
