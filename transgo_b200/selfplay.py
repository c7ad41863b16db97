"""Batched replacement of the per-actor self-play loop (self_play.py: SelfPlay.continuous_self_play, :902-968):
G games advance together on one GPU; positions, search policies and players are kept on the device, labelled when a game
ends (winner z, ownership) and expanded with the 8 board symmetries exactly as the reference does before `mem.append`.
"""
import torch

from .mcts import BatchedMCTS


def augment8(obs, pi, z, own, n):
    """The reference's x8 augmentation (self_play.py:943-965) for a batch: for i in 1..4: rot90 by i, then its fliplr.
    obs [B,10,n,n], pi [B,A], z [B], own [B,P] -> tensors with 8B rows ordered (position, i, rot|flip) like the loop."""
    B = obs.shape[0]
    outs_o, outs_p, outs_w = [], [], []
    board_pi = pi[:, :-1].reshape(B, n, n)
    board_own = own.reshape(B, n, n)
    for i in (1, 2, 3, 4):
        ro, rp, rw = torch.rot90(obs, i, (2, 3)), torch.rot90(board_pi, i, (1, 2)), torch.rot90(board_own, i, (1, 2))
        outs_o += [ro, torch.flip(ro, (3,))]
        outs_p += [rp, torch.flip(rp, (2,))]
        outs_w += [rw, torch.flip(rw, (2,))]
    o = torch.stack(outs_o, 1).reshape(B * 8, *obs.shape[1:])
    p = torch.stack(outs_p, 1).reshape(B * 8, n * n)
    p = torch.cat([p, pi[:, -1:].repeat_interleave(8, 0)], 1)
    w = torch.stack(outs_w, 1).reshape(B * 8, n * n)
    return o, p, z.repeat_interleave(8, 0), w


def assemble_trajectories(hist_obs, hist_pi, hist_player, length, winner, terr, fin=None, augment=True, out=None, index=0):
    """self_play.py:931-965 for a batch of games in one kernel launch (csrc/tg_traj.cu): label every recorded position of the
    finished games (z from the winner, ownership from the mover's view), expand x8 in the reference's symmetry order and write
    the rows either into fresh tensors (out=None) or straight into a replay ring's tensors at its write position
    (out=(obs, pi, z, own), index=ring.index; rows wrap modulo the capacity).
    hist_obs [G,T,C,n,n] f32, hist_pi [G,T,A] f32, hist_player [G,T] i32, length [G] i64, winner [G] u8, terr [G,P] f32,
    fin [G] bool/u8 or None.  Returns (obs, pi, z, own, rows)."""
    from . import _lib
    L = _lib.lib()
    G, T, C, n = hist_obs.shape[0], hist_obs.shape[1], hist_obs.shape[2], hist_obs.shape[3]
    A, P = n * n + 1, n * n
    dev = hist_obs.device
    sym = 8 if augment else 1
    length = None if length is None else length.to(torch.int64).contiguous()
    fin_u8 = None if fin is None else fin.to(torch.uint8).contiguous()
    if length is None:
        rows = (G if fin is None else int(fin_u8.sum().item())) * sym
    else:
        rows = int((length if fin is None else length * fin_u8.to(torch.int64)).sum().item()) * sym
    if out is None:
        cap = max(rows, 1)
        out = (torch.empty([cap, C, n, n], dtype=torch.float32, device=dev), torch.empty([cap, A], dtype=torch.float32, device=dev),
               torch.empty([cap], dtype=torch.float32, device=dev), torch.empty([cap, P], dtype=torch.float32, device=dev))
        index = 0
    o_obs, o_pi, o_z, o_own = out
    cap = o_obs.shape[0]
    assert rows <= cap, "more rows than the output holds: append in smaller batches"
    assert all(t.is_contiguous() and t.dtype == torch.float32 for t in (hist_obs, hist_pi, terr, o_obs, o_pi, o_z, o_own))
    _lib.check(L.tg_traj_assemble(n, C, G, T, _lib.ptr(hist_obs), _lib.ptr(hist_pi), _lib.ptr(hist_player.to(torch.int32).contiguous()),
                                  _lib.ptr(length), _lib.ptr(fin_u8), _lib.ptr(winner.to(torch.uint8).contiguous()), _lib.ptr(terr),
                                  _lib.ptr(o_obs), _lib.ptr(o_pi), _lib.ptr(o_z), _lib.ptr(o_own), cap, int(index), sym, None,
                                  _lib.stream_ptr(device=dev)), "tg_traj_assemble")
    if cap == max(rows, 1) and index == 0:
        return o_obs[:rows], o_pi[:rows], o_z[:rows], o_own[:rows], rows
    return o_obs, o_pi, o_z, o_own, rows


def augment8_rows(obs, pi, z, own, n, out=None, index=0):
    """The x8 expansion of ALREADY LABELLED rows (the receiver side of selfplay_round: the wire carries each position once) with
    the same kernel: every row is a one-position game whose mover is black and whose winner is black iff z > 0, so the kernel's
    labels reproduce z and own.  Writes into fresh tensors or into a ring's tensors at `index`."""
    B = obs.shape[0]
    dev = obs.device
    if B == 0:
        return (obs, pi, z, own) if out is None else out
    player = torch.ones([B, 1], dtype=torch.int32, device=dev)
    winner = torch.where(z > 0, 1, 2).to(torch.uint8)
    o, p, zz, w, _ = assemble_trajectories(obs.reshape(B, 1, *obs.shape[1:]).float().contiguous(), pi.reshape(B, 1, -1).float().contiguous(),
                                           player, None, winner, own.float().contiguous(), augment=True, out=out, index=index)
    return o, p, zz, w


class SelfPlayWorker:
    """One self-play rank: plays `num_games` games to the end and returns the labelled (optionally augmented) positions."""

    def __init__(self, evaluator, num_games, board_size=9, device="cuda", **mcts_kw):
        self.n = board_size
        self.search = BatchedMCTS(evaluator, num_games, board_size=board_size, device=device, **mcts_kw)

    def play_games(self, augment=True, max_moves=None):
        s = self.search
        G, n = s.G, self.n
        s.reset_root()
        obs_l, pi_l, player_l, alive_l = [], [], [], []
        alive = torch.ones([G], dtype=torch.bool, device=s.device)
        moves = 0
        while bool(alive.any()) and (max_moves is None or moves < max_moves):
            actions, pi, obs = s.get_action_probs(is_selfplay=True)
            player = 1 + ((s.step - 1) % 2)                  # black moves on odd step counts (board.cc:15,19,537)
            obs_l.append(obs.clone()); pi_l.append(pi.to(torch.float32)); player_l.append(player.clone()); alive_l.append(alive.clone())
            done = s.update_with_action(actions)
            alive = alive & (done == 0)
            moves += 1
        winner, score, terr = s.outcome(want_territory=True)          # self_play.py:931-937
        obs = torch.stack(obs_l, 1).contiguous(); pi = torch.stack(pi_l, 1).contiguous()   # [G, T, ...]
        player = torch.stack(player_l, 1); valid = torch.stack(alive_l, 1)
        per_game = valid.sum(1)                                       # games stay alive for a prefix of the moves
        obs, pi, z, own, rows = assemble_trajectories(obs, pi, player, per_game, winner, terr, augment=augment)
        return dict(obs=obs, pi=pi, z=z, own=own, winner=winner, score=score, moves=moves,
                    positions=rows // (8 if augment else 1), positions_per_game=per_game)


def _play_continuous(self, min_positions, augment=True, max_moves=None, pacing=None, wait_fn=None, game_total_num=None):
    """continuous_self_play (self_play.py:902-968) for a batch: every slot plays game after game; a slot whose game ends is
    labelled (winner z and ownership from the mover's view), handed out and restarted at once, so the search batch stays
    full while games end at different plies.  Returns at least `min_positions` labelled positions (x8 when `augment`) of
    FINISHED games; unfinished games stay in their slots for the next call.

    `pacing` (replay.Pacing) gets the reference's counter updates per finished game and throttles the loop like
    self_play.py:970-980: while the train/play ratio is below target, `wait_fn()` is called (it lets the trainer catch up:
    a train step, a counter refresh from the trainer rank, or a sleep); without a `wait_fn` the call returns early with
    throttled=True so that the caller can train."""
    s = self.search
    G, n, A = s.G, self.n, s.A
    T = int(s.cfg.max_step) + 1
    dev = s.device
    if getattr(self, "_hist", None) is None:
        s.reset_root()
        self._hist = dict(obs=torch.zeros([G, T, 10, n, n], dtype=torch.float32, device=dev),
                          pi=torch.zeros([G, T, A], dtype=torch.float32, device=dev),
                          player=torch.zeros([G, T], dtype=torch.int32, device=dev),
                          length=torch.zeros([G], dtype=torch.int64, device=dev))
    h = self._hist
    gi = torch.arange(G, device=dev)
    out = dict(obs=[], pi=[], z=[], own=[])
    positions = games = moves = 0
    throttled = False
    while positions < min_positions and (max_moves is None or moves < max_moves):
        actions, pi, obs = s.get_action_probs(is_selfplay=True)
        t = h["length"].clamp_max(T - 1)
        h["obs"][gi, t] = obs
        h["pi"][gi, t] = pi.to(torch.float32)
        h["player"][gi, t] = (1 + ((s.step - 1) % 2)).to(torch.int32)            # black moves on odd step counts (board.cc:15,19,537)
        h["length"] += 1
        fin = s.update_with_action(actions).bool()
        moves += 1
        if bool(fin.any()):
            winner, _, terr = s.outcome(want_territory=True)       # self_play.py:931-937 (roots of finished games are final)
            o, p, zz, w, rows = assemble_trajectories(h["obs"], h["pi"], h["player"], h["length"], winner, terr, fin=fin, augment=augment)
            out["obs"].append(o); out["pi"].append(p); out["z"].append(zz); out["own"].append(w)
            positions += rows // (8 if augment else 1)
            games += int(fin.sum().item())
            if pacing is not None:
                for npos in h["length"][fin].tolist():
                    pacing.finish_game(int(npos))
            h["length"][fin] = 0
            s.reset_games(fin)
            if pacing is not None and pacing.selfplay_should_wait(game_total_num):
                if wait_fn is None:
                    throttled = True
                    break
                while pacing.selfplay_should_wait(game_total_num):
                    wait_fn()
    if out["obs"]:
        obs, pi, z, own = (torch.cat(out[k]) for k in ("obs", "pi", "z", "own"))
    else:
        obs, pi = torch.zeros([0, 10, n, n], device=dev), torch.zeros([0, A], device=dev)
        z, own = torch.zeros([0], device=dev), torch.zeros([0, n * n], device=dev)
    return dict(obs=obs, pi=pi, z=z, own=own, positions=positions, games=games, moves=moves, throttled=throttled)


SelfPlayWorker.play_continuous = _play_continuous


class _Stopwatch:
    """Phase timer for selfplay_round: CUDA events on the current stream when a GPU is in use, wall clock otherwise."""

    def __init__(self, cuda):
        self.cuda, self.marks = cuda, []

    def mark(self, name):
        if self.cuda:
            e = torch.cuda.Event(enable_timing=True); e.record()
        else:
            import time
            e = time.perf_counter()
        self.marks.append((name, e))

    def phases(self):
        if self.cuda:
            torch.cuda.synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.marks, self.marks[1:]):
            out[name + "_ms"] = a.elapsed_time(b) if self.cuda else (b - a) * 1e3
        return out


def selfplay_round(worker, model, ring=None, pacing=None, src=0, dst=0, augment=True, max_moves=None, refresh=None, timing=False):
    """One round of BASELINE configs[4] on every rank of the process group (or a single process without one):

        weights  trainer rank `src` -> all ranks            (replaces shared_storage.get_info("weights"), self_play.py:913)
        play     this rank's shard of games to the end      (WP_MCTS loop, self_play.py:915-929; no cross-GPU traffic)
        gather   labelled positions -> rank `dst`           (replaces mem.append.remote, self_play.py:956-965)
        augment  x8 on the receiver (self_play.py:943-965), so the wire carries each position once
        append   into the replay ring on `dst`, bump the play counters (shared_storage.py:21-43)

    `refresh` re-folds inference copies of the weights (e.g. make_fused_evaluator(...).refresh).  Returns the play
    statistics of this rank plus, on `dst`, the number of rows appended; with `timing`, also broadcast_ms / play_ms /
    gather_ms / append_ms of this rank (device time) and the bytes each collective moved."""
    import torch.distributed as dist
    from .dist import broadcast_weights, gather_trajectories
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    sw = _Stopwatch(next(model.parameters()).is_cuda) if timing else None
    if sw: sw.mark("start")
    if multi:
        broadcast_weights(model, src=src)
    if refresh is not None:
        refresh()
    if sw: sw.mark("broadcast")
    out = worker.play_games(augment=False, max_moves=max_moves)
    rows = (out["obs"], out["pi"], out["z"], out["own"])
    games_here = int(out["winner"].numel())
    per_game = out["positions_per_game"]                       # int64 [G] on the device: positions of each game of this rank
    if sw: sw.mark("play")
    if multi:
        # the row-count exchange also carries every rank's game count and per-game position counts: shards are uneven when
        # the total is not divisible by the world size, and the pacing schedules count games one by one
        rows, extras = gather_trajectories(*rows, dst=dst, extra=[games_here])
        games_all = [int(x) for x in extras[:, 0].tolist()]
    else:
        games_all = [games_here]
    if sw: sw.mark("gather")
    rank = dist.get_rank() if multi else dst
    if augment and rank == dst and rows is not None:
        rows = augment8_rows(*rows, worker.n)
    stats = dict(moves=out["moves"], positions=out["positions"], games=games_here, games_all_ranks=int(sum(games_all)),
                 black_wins=int((out["winner"] == 1).sum().item()), appended=0)
    if rank == dst and rows is not None:
        if ring is not None:
            ring.append(*rows)
        stats["appended"] = int(rows[0].shape[0])
        if pacing is not None:
            # counters advance game by game in the reference's order (positions of a game, then the game itself:
            # self_play.py:921, :967); only the totals of the other ranks are known here, spread evenly over their games
            total_pos = stats["appended"] // (8 if augment else 1)
            total_games = int(sum(games_all))
            local = per_game.tolist() if not multi else None
            for i in range(total_games):
                if local is not None:
                    pacing.finish_game(int(local[i]))
                else:
                    pacing.finish_game(total_pos // total_games + (1 if i < total_pos % total_games else 0))
    if sw:
        sw.mark("append")
        stats.update(sw.phases())
        row_bytes = 4 * (out["obs"][0].numel() + out["pi"].shape[1] + 1 + out["own"].shape[1]) if out["positions"] else 0
        stats["gather_bytes_sent"] = 0 if (not multi or rank == dst) else out["positions"] * row_bytes
        stats["broadcast_bytes"] = sum(v.numel() * v.element_size() for v in model.state_dict().values()) if multi else 0
    return stats


def _root_state_view(search):
    """int32 [SQ, G, 4] view of the engine's root states inside its workspace (same layout as BatchedGoEnv.states)."""
    sq = search.lib.tg_state_bytes(search.n) // 16
    base = search.lib.tg_mcts_root_states(search._cfg, search._w())
    off = (base - search.ws.data_ptr()) // 4
    return search.ws.reshape(-1)[off:off + sq * search.G * 4].reshape(sq, search.G, 4)


def evaluate_match(evaluator_new, evaluator_old, n_games=10, board_size=9, device="cuda", seed=0, return_actions=False, **mcts_kw):
    """Batched form of SelfPlay.policy_evaluate (self_play.py:987-1040): `n_games` games between the newest weights and the
    "evaluate" weights, colours alternating from game to game (:1000, :1026), every move chosen by a FRESH search of the side
    to move from the current position (select_action, :689-704: no tree reuse, no root noise, temperature 0.12).

    All games advance together.  At every ply the new net searches the games in which it is to move and the old net the
    others, each in its own engine of n_games / 2 slots.  Returns dict(win, lose, win_ratio, winners, new_plays_black)."""
    from . import _lib
    from .environment import BatchedGoEnv
    assert n_games % 2 == 0, "colours alternate: use an even number of games"
    G, half = int(n_games), int(n_games) // 2
    env = BatchedGoEnv(G, board_size=board_size, device=device, max_step=mcts_kw.get("max_step"), komi=mcts_kw.get("komi", 7.5))
    eng_new = BatchedMCTS(evaluator_new, half, board_size=board_size, device=device, **mcts_kw)
    eng_old = BatchedMCTS(evaluator_old, half, board_size=board_size, device=device, **mcts_kw)
    dev = env.device
    new_black = torch.arange(G, device=dev) % 2 == 0                   # game i: the new net holds BLACK iff i is even
    idx_a, idx_b = torch.nonzero(new_black).reshape(-1), torch.nonzero(~new_black).reshape(-1)
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    fresh = BatchedGoEnv(1, board_size=board_size, device=device).states      # an empty board for finished slots
    ply = 0
    played = []
    while not bool(env.done.all()) and ply <= env.max_step + 1:
        black_to_move = ply % 2 == 0                                      # every game alternates strictly (a pass is a move)
        actions = torch.zeros([G], dtype=torch.int16, device=dev)
        for eng, idx in ((eng_new, idx_a if black_to_move else idx_b), (eng_old, idx_b if black_to_move else idx_a)):
            _lib.check(eng.lib.tg_mcts_init(eng._cfg, eng._w(), eng._s()), "tg_mcts_init")
            st = env.states[:, idx]
            over = env.done[idx].bool()
            st = torch.where(over[None, :, None], fresh.expand(-1, half, -1), st)
            _root_state_view(eng).copy_(st)
            eng._expand_roots()
            u = torch.rand([half], dtype=torch.float64, device=dev, generator=gen)
            a, _, _ = eng.get_action_probs(is_selfplay=False, u=u)
            actions[idx] = a
        if return_actions:
            played.append(actions.clone())
        env.step(actions, want_legal=False)
        ply += 1
    score = env.score()
    winners = torch.where(score > 0, 1, 2)                                # environment.py:118-119
    new_won = torch.where(new_black, winners == 1, winners == 2)
    win = int(new_won.sum().item())
    out = dict(win=win, lose=G - win, win_ratio=win / G, winners=winners, new_plays_black=new_black, plies=ply)
    if return_actions:
        out["actions"] = torch.stack(played) if played else torch.zeros([0, G], dtype=torch.int16, device=dev)   # [plies, G]; finished games: ignored
    return out
