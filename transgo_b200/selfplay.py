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
        obs = torch.stack(obs_l, 1); pi = torch.stack(pi_l, 1)         # [G, T, ...]
        player = torch.stack(player_l, 1); valid = torch.stack(alive_l, 1)
        z = torch.where(player == winner[:, None].to(player.dtype), 1.0, -1.0)
        own = torch.where((player == 1)[:, :, None], terr[:, None, :], -terr[:, None, :])
        obs, pi, z, own = obs[valid], pi[valid], z[valid], own[valid]
        if augment:
            obs, pi, z, own = augment8(obs, pi, z, own, n)
        return dict(obs=obs, pi=pi, z=z, own=own, winner=winner, score=score, moves=moves,
                    positions=int(valid.sum().item()))
