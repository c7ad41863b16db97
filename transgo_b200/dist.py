"""Off-hot-path collectives: one process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests).  Games shard by global id and never communicate while playing; only two exchanges exist:

* broadcast_weights   trainer rank -> self-play ranks; replaces `shared_storage.get_info("weights")` (self_play.py:913,
                      trainer.py:75-79).  The 226 state_dict tensors travel as ONE flat buffer per dtype (11.9 MB fp32).
* gather_trajectories finished games -> replay owner; replaces `mem.append.remote` (self_play.py:956-965).  Positions are
                      concatenated, lengths exchanged first, payload padded to the longest rank.
"""
import torch
import torch.distributed as dist


def shard_games(total_games, rank, world):
    """(game_id0, count) of this rank's contiguous slice; identical games for any world size (hash-keyed randomness)."""
    base, rem = divmod(int(total_games), int(world))
    count = base + (1 if rank < rem else 0)
    return rank * base + min(rank, rem), count


def _buckets(state_dict):
    by_dtype = {}
    for name in sorted(state_dict.keys()):
        t = state_dict[name]
        by_dtype.setdefault(t.dtype, []).append((name, t))
    return by_dtype


def broadcast_weights(model, src=0, group=None, device=None):
    """In-place: after the call every rank's model holds rank `src`'s parameters and buffers."""
    sd = model.state_dict()
    for dtype, items in _buckets(sd).items():
        dev = device or items[0][1].device
        flat = torch.cat([t.detach().reshape(-1).to(dev) for _, t in items])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for _, t in items:
            n = t.numel()
            t.copy_(flat[off:off + n].reshape(t.shape))
            off += n
    return model


def gather_trajectories(obs, pi, z, own, dst=0, group=None):
    """obs [n_i,10,b,b] f32, pi [n_i,A], z [n_i], own [n_i,P] on every rank -> concatenation over ranks on `dst`
    (None elsewhere).  n_i may differ per rank (games end at different plies)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = obs.device
    n = torch.tensor([obs.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    shape = list(obs.shape[1:])
    no, na, nown = int(torch.tensor(shape).prod().item()), pi.shape[1], own.shape[1]
    width = no + na + 1 + nown
    packed = torch.zeros([m, width], dtype=torch.float32, device=dev)
    if obs.shape[0]:
        packed[:obs.shape[0]] = torch.cat([obs.reshape(obs.shape[0], -1).float(), pi.float(), z.float().reshape(-1, 1), own.float()], 1)
    out = [torch.zeros_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed, out, dst=dst, group=group)
    if rank != dst:
        return None
    allp = torch.cat([o[:c] for o, c in zip(out, counts)], 0)
    return (allp[:, :no].reshape([-1] + shape), allp[:, no:no + na], allp[:, no + na], allp[:, no + na + 1:no + na + 1 + nown])
