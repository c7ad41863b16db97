"""Off-hot-path collectives: one process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests).  Games shard by global id and never communicate while playing; only two exchanges exist:

* broadcast_weights   trainer rank -> self-play ranks; replaces `shared_storage.get_info("weights")` (self_play.py:913,
                      trainer.py:75-79).  The 226 state_dict tensors travel as ONE flat buffer per dtype (11.9 MB fp32).
* gather_trajectories finished games -> replay owner; replaces `mem.append.remote` (self_play.py:956-965).  Row counts are
                      exchanged with one all_gather_into_tensor, then each rank sends its rows as one variable-size message.
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


def gather_trajectories(obs, pi, z, own, dst=0, group=None, extra=None):
    """obs [n_i,10,b,b] f32, pi [n_i,A], z [n_i], own [n_i,P] on every rank -> concatenation over ranks on `dst`
    (None elsewhere).  n_i may differ per rank (games end at different plies).

    One `all_gather_into_tensor` carries every rank's row count (and `extra`, a small int64 vector per rank such as its
    number of finished games), one host read sizes the receive buffers, and the payload travels as ONE variable-size
    point-to-point message per rank (batch_isend_irecv: NCCL send/recv over NVLink, gloo on CPU) -- nothing is padded to
    the longest rank.  Returns (obs, pi, z, own) on dst; with `extra`, returns ((obs, pi, z, own) | None, extras [world, len(extra)])
    on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = obs.device
    ex = torch.zeros([0], dtype=torch.int64, device=dev) if extra is None else torch.as_tensor(extra, dtype=torch.int64, device=dev).reshape(-1)
    mine = torch.cat([torch.tensor([obs.shape[0]], dtype=torch.int64, device=dev), ex])
    table = torch.empty([world * mine.numel()], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(table, mine, group=group)
    table = table.reshape(world, -1).cpu()                      # the one host read
    counts = table[:, 0].tolist()
    shape = list(obs.shape[1:])
    no = 1
    for d in shape:
        no *= int(d)
    na, nown = pi.shape[1], own.shape[1]
    width = no + na + 1 + nown
    n = obs.shape[0]
    packed = torch.cat([obs.reshape(n, -1).float(), pi.float(), z.float().reshape(-1, 1), own.float()], 1) if n else \
        torch.zeros([0, width], dtype=torch.float32, device=dev)
    result = None
    if rank == dst:
        total = int(sum(counts))
        allp = torch.empty([total, width], dtype=torch.float32, device=dev)
        ops, off = [], 0
        for r, c in enumerate(counts):
            if c and r != dst:
                ops.append(dist.P2POp(dist.irecv, allp[off:off + c], r if group is None else dist.get_global_rank(group, r), group))
            elif c:
                allp[off:off + c] = packed
            off += c
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        result = (allp[:, :no].reshape([-1] + shape), allp[:, no:no + na], allp[:, no + na], allp[:, no + na + 1:no + na + 1 + nown])
    elif n:
        peer = dst if group is None else dist.get_global_rank(group, dst)
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, packed, peer, group)]):
            w.wait()
    if extra is None:
        return result
    return result, table[:, 1:]
