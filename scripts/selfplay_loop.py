#!/usr/bin/env python
"""BASELINE configs[4] end to end: sharded self-play with NCCL weight broadcast and trajectory gather.

    python scripts/selfplay_loop.py [--games 256 --sims 64 --rounds 2 --max-moves 12]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/selfplay_loop.py ...

Every rank plays `games` games per round with the batched search and the fused TransGo evaluator; rank 0 owns the replay
ring (and would host the trainer).  Prints one JSON line per round on rank 0 (positions/s, sims/s, ring fill)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=256)
    ap.add_argument("--sims", type=int, default=64)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--max-moves", type=int, default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from transgo_b200.dist import shard_games
    from transgo_b200.fused_net import make_fused_evaluator
    from transgo_b200.net import TransGoNetwork
    from transgo_b200.replay import Pacing, ReplayRing
    from transgo_b200.selfplay import SelfPlayWorker, selfplay_round
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)                       # ranks start with DIFFERENT weights; the broadcast must fix that
    model = TransGoNetwork(board_size=9).to(dev).eval()
    ev = make_fused_evaluator(model, cuda_graph=True)
    game_id0, _ = shard_games(args.games * world, rank, world)
    worker = SelfPlayWorker(ev, args.games, num_simulation=args.sims, tie_seed=7, game_id0=game_id0, device=dev)
    ring = ReplayRing(1 << 18, device=dev) if rank == 0 else None
    pacing = Pacing() if rank == 0 else None
    for r in range(args.rounds):
        if rank == 0:
            with torch.no_grad():                        # stand-in for a train step: the trainer rank's weights change
                for p in model.parameters():
                    p.mul_(1.0 + 1e-3 * (r + 1))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a.record()
        worker.search.total_sims()
        st = selfplay_round(worker, model, ring, pacing, src=0, dst=0, max_moves=args.max_moves, refresh=ev.refresh)
        if rank == 0:
            # the throttle of self_play.py:970-980 / trainer.py:94-105: self-play waits while the trainer is behind the
            # train/play ratio; here the "trainer" is the stand-in above, so it just catches up
            while pacing.selfplay_should_wait():
                pacing.now_train_steps += 1
        sims = worker.search.total_sims()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        tot = torch.tensor([sims, st["positions"]], dtype=torch.float64, device=dev)
        digest = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        lo, hi = digest.clone(), digest.clone()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"round": r, "n_gpus": world, "games": args.games * world, "moves": st["moves"], "ms": float(ms[0]),
                              "mcts_sims_per_s": float(tot[0]) / float(ms[0]) * 1e3, "positions": int(tot[1]),
                              "positions_per_s": float(tot[1]) / float(ms[0]) * 1e3, "rows_appended_x8": st["appended"],
                              "ring_fill": len(ring), "weights_identical_on_all_ranks": bool(float(lo[0]) == float(hi[0])),
                              "now_play_steps": pacing.now_play_steps, "now_play_games": pacing.now_play_games}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
