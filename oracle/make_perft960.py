"""Writes tests/golden/perft960.npz: perft(1..5) of all 960 Chess960 start positions (chess.Board.from_chess960_pos(id),
castling in Chess960 mode) and of the vanilla start position, counted by the oracle's C rules engine (oracle/chess_oracle.c).
BASELINE config c1.  Test infrastructure: the GPU engine and the host build of chess.cuh are compared with these numbers.

    python -m oracle.make_perft960 [workers]
"""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEPTH = 5


def perft_row(sid):
    """sid 0..959: Chess960 start position; -1: chess.Board() (vanilla castling rules)"""
    from oracle import ref_path  # noqa: F401  (puts the stand-in `chess` on sys.path)
    import chess
    b = chess.Board() if sid < 0 else chess.Board.from_chess960_pos(sid)
    return [b.perft(d) for d in range(1, DEPTH + 1)]


def main():
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    ids = list(range(960)) + [-1]
    with ProcessPoolExecutor(workers) as ex:
        rows = list(ex.map(perft_row, ids, chunksize=8))
    nodes = np.array(rows, dtype=np.uint64)
    assert nodes[518].tolist() == [20, 400, 8902, 197281, 4865609] and nodes[960].tolist() == nodes[518].tolist()
    out = os.path.join(ROOT, "tests", "golden", "perft960.npz")
    np.savez_compressed(out, ids=np.array(ids, dtype=np.int16), nodes=nodes)
    print("wrote", out, nodes.shape, "sum depth-5 nodes", int(nodes[:, 4].sum()))


if __name__ == "__main__":
    main()
