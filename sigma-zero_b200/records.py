"""Training-data record format for self-play output: what sim.selfplay_records writes and what the library's trainer
(trainer.py -> szb_train_records) reads.

The reference stores one uncompressed bool[119,8,8] tensor and one {Move: probability} dict per position
(sim.py:56,71-72) and train_RL.chessDataset / collatefn (:14-49) expect bit-packed states.  Here a batch of game
histories (the dicts sim.selfplay_batch / play_game return) becomes flat arrays:

    states   uint64[N, 119]   one word per input plane, bit (row*8+col) -- exactly what szb_encode emits
    pi_index uint16[M]        policy indices (chess_tensor.py:221-306 layout) of the searched root children, CSR
    pi_prob  float32[M]       their visit fractions (mcts.py:113-122)
    pi_off   int64[N+1]       CSR offsets: position i owns pi_index[pi_off[i]:pi_off[i+1]]
    z        int8[N]          game outcome from the mover's point of view (sim.py:86-97)
    colour   bool[N]          side to move (True = White)
    game     int32[N]         game number inside the batch

952 bytes per position for the state instead of 7,616.  Host logic only."""
import numpy as np

from . import runtime
from .chess_tensor import move_index


def pack_records(games):
    states, idx, prob, off, z, colour, game = [], [], [], [0], [], [], []
    for g, h in enumerate(games):
        for s, a, r, c in zip(h["states"], h["actions"], h["rewards"], h["colours"]):
            states.append(runtime.pack_planes(np.asarray(s)))
            for m, p in a.items():
                idx.append(move_index(m, bool(c)))
                prob.append(p)
            off.append(len(idx))
            z.append(r)
            colour.append(bool(c))
            game.append(g)
    return {"states": np.array(states, dtype=np.uint64).reshape(-1, 119), "pi_index": np.array(idx, dtype=np.uint16),
            "pi_prob": np.array(prob, dtype=np.float32), "pi_off": np.array(off, dtype=np.int64),
            "z": np.array(z, dtype=np.int8), "colour": np.array(colour, dtype=bool), "game": np.array(game, dtype=np.int32)}


def unpack_states(records, rows=None):
    """bool[n,119,8,8] planes of the selected rows (all by default)"""
    s = records["states"] if rows is None else records["states"][rows]
    return runtime.unpack_planes(s)


def dense_policy(records, rows=None):
    """float32[n,4672] soft targets (train_RL.collatefn's actionsToTensor on the stored dicts)"""
    rows = range(len(records["z"])) if rows is None else rows
    out = np.zeros((len(rows), 4672), dtype=np.float32)
    for k, i in enumerate(rows):
        lo, hi = records["pi_off"][i], records["pi_off"][i + 1]
        out[k, records["pi_index"][lo:hi]] = records["pi_prob"][lo:hi]
    return out


def save(path, records):
    np.savez_compressed(path, **records)


def load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}
