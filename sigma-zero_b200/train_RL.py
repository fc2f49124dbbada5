"""Self-play fan-out of the reference's train_RL.py (args + launcher only, train_RL.py:156-244).

The reference spawns `num_process` CPU workers that each play `num_games // num_process` games one after another
and merges their pickled dicts.  Here every rank (one process per GPU, torch.distributed over NCCL) plays its
shard of the games concurrently on its GPU; the only collective is the broadcast of the flat fp32 weight buffer
from the trainer rank at the start of an iteration.  `num_selfPlay_iterations` -- which the reference's args
carry but never read -- is defined as the number of self-play games per outer iteration.
Training itself (chessDataset / train / test) is outside the self-play hot path."""
import os

import torch

DEFAULT_ARGS = {
    'C': 2,
    'num_searches': 100,
    'num_iterations': 3,
    'num_selfPlay_iterations': 500,
    'num_epochs': 4,
    'batch_size': 64,
    'chess960': True,
}


def shard_of(n_games, rank, world):
    """contiguous block of game indices owned by `rank` (sizes differ by at most one)"""
    base, extra = divmod(n_games, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def flatten_state_dict(sd):
    keys = [k for k in sd if not k.endswith("num_batches_tracked")]
    return keys, torch.cat([sd[k].detach().reshape(-1).float() for k in keys])


def unflatten_into(model, keys, flat):
    sd = model.state_dict()
    off = 0
    for k in keys:
        n = sd[k].numel()
        sd[k].copy_(flat[off:off + n].reshape(sd[k].shape))
        off += n


def broadcast_weights(model, src=0, device=None):
    """one flat-buffer broadcast (NCCL over NVLink on GPUs, gloo in CPU tests); returns the buffer's bytes"""
    import torch.distributed as dist
    keys, flat = flatten_state_dict(model.state_dict())
    if device is not None:
        flat = flat.to(device)
    dist.broadcast(flat, src=src)
    unflatten_into(model, keys, flat.cpu())
    return flat.numel() * 4


def selfplay_iteration(model, args, seed=0, max_plies=None):
    """One outer iteration's self-play: this rank's shard of args['num_selfPlay_iterations'] games."""
    import torch.distributed as dist
    from .sim import selfplay_batch
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if dist.is_initialized() and world > 1:
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if torch.cuda.is_available() else None
        broadcast_weights(model, 0, dev)
    lo, hi = shard_of(int(args['num_selfPlay_iterations']), rank, world)
    games, counters = selfplay_batch(model, args, hi - lo, c960=bool(args.get('chess960', False)),
                                     seed=seed, max_plies=max_plies, game_id_base=lo)
    return games, counters
