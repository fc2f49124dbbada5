"""Shared helpers for the parity tests (test infrastructure)."""
import ctypes
import os
import subprocess

import numpy as np

from oracle import hash_eval, ref_path  # noqa: F401  (puts oracle/ on sys.path so `chess` is the stand-in)
import chess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_game(c960, sid, moves=()):
    g = ref_path.RefGame(chess960=bool(c960), start_id=int(sid) if c960 else None)
    for u in moves:
        g.move_piece(chess.Move.from_uci(u))
    return g


def wire_pos(board, lib_mod):
    """oracle Board -> szb_pos"""
    bb, turn, rw, rb, ep, hm, ply = board._export()
    p = lib_mod.Pos()
    for i in range(12):
        p.pieces[i] = bb[i]
    p.turn, p.castling_w, p.castling_b, p.ep_square = turn, rw, rb, ep
    p.halfmove_clock, p.ply, p.chess960 = hm, ply, int(board.chess960)
    return p


def setup_games(engine, specs):
    """specs: list of (c960, sid, [uci moves]).  Replays them on the GPU through szb_games_push and returns the
    oracle RefGame objects in the same order."""
    engine.reset([int(s) if c else -1 for c, s, _ in specs])
    games = [oracle_game(c, s) for c, s, _ in specs]
    longest = max(len(m) for _, _, m in specs)
    for ply in range(longest):
        who, idx = [], []
        for g, (_, _, moves) in enumerate(specs):
            if ply < len(moves):
                mv = chess.Move.from_uci(moves[ply])
                who.append(g)
                idx.append(ref_path.move_to_index(mv, games[g].board.turn))
                games[g].move_piece(mv)
        engine.push(who, idx)
    return games


def legal_indices(game):
    b = game.board
    return sorted(ref_path.move_to_index(m, b.turn) for m in b.legal_moves)


def outcome_code(board):
    o = board.outcome()
    return o.termination if o else 0


def build_host_harness():
    src = os.path.join(ROOT, "tests", "host_harness", "harness.cpp")
    out = os.path.join(ROOT, "tests", "host_harness", "_harness.so")
    deps = [src] + [os.path.join(ROOT, "sigma-zero_b200", "csrc", f) for f in ("chess.cuh", "tree.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-o", out, src])
    H = ctypes.CDLL(out)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    H.hh_new.restype = vp; H.hh_new.argtypes = [ci]
    H.hh_set.restype = vp; H.hh_set.argtypes = [vp] + [ci] * 7
    H.hh_free.argtypes = [vp]
    H.hh_perft.restype = ctypes.c_uint64; H.hh_perft.argtypes = [vp, ci]
    for f in ("hh_legal", "hh_push_index", "hh_outcome", "hh_flags", "hh_codec_roundtrip", "hh_turn"):
        getattr(H, f).restype = ci
    H.hh_legal.argtypes = [vp, vp]; H.hh_push_index.argtypes = [vp, ci]; H.hh_planes.argtypes = [vp, vp]
    H.hh_outcome.argtypes = [vp]; H.hh_flags.argtypes = [vp]; H.hh_codec_roundtrip.argtypes = [vp]; H.hh_turn.argtypes = [vp]
    H.hh_puct.restype = ctypes.c_float
    H.hh_puct.argtypes = [ci, ctypes.c_double, ctypes.c_float, ci, ctypes.c_float]
    H.hh_cascade_sum.restype = ctypes.c_float; H.hh_cascade_sum.argtypes = [vp]
    H.hh_cascade_sum_sparse.restype = ctypes.c_float; H.hh_cascade_sum_sparse.argtypes = [vp, vp]
    H.hh_noisy_prior.restype = ctypes.c_float; H.hh_noisy_prior.argtypes = [ctypes.c_float]
    H.hh_hash_eval.argtypes = [vp, vp, vp]
    return H


PERFT_KATS = [
    ("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", False, [20, 400, 8902, 197281, 4865609, 119060324]),
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", False, [48, 2039, 97862, 4085603, 193690690]),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", False, [14, 191, 2812, 43238, 674624]),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", False, [6, 264, 9467, 422333, 15833292]),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", False, [44, 1486, 62379, 2103487, 89941194]),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", False, [46, 2079, 89890, 3894594, 164075551]),
    ("bqnb1rkr/pp3ppp/3ppn2/2p5/5P2/P2P4/NPP1P1PP/BQ1BNRKR w HFhf - 2 9", True, [21, 528, 12189, 326672, 8146062]),
    ("2nnrbkr/p1qppppp/8/1ppb4/6PP/3PP3/PPP2P2/BQNNRBKR w HEhe - 1 9", True, [21, 807, 18002, 667366, 16253601]),
    ("b1q1rrkb/pppppppp/3nn3/8/P7/1PPP4/4PPPP/BQNNRKRB w GE - 1 9", True, [20, 479, 10471, 273318, 6417013]),
]


PGN_SAMPLE = """[Event "Opera"]
[Site "Paris"]
[Result "1-0"]
[PlyCount "33"]

1.e4 e5 2.Nf3 d6 3.d4 Bg4 {a pin} 4.dxe5 Bxf3 5.Qxf3 dxe5 6.Bc4 Nf6 7.Qb3 Qe7 8.Nc3 c6 9.Bg5 b5 $2 10.Nxb5 cxb5
11.Bxb5+ Nbd7 12.O-O-O Rd8 (12...Qb4 13.Bxf6) 13.Rxd7 Rxd7 14.Rd1 Qe6 15.Bxd7+ Nxd7 16.Qb8+ Nxb8 17.Rd8# 1-0

[Event "en passant, under-promotion by capture, short castling"]
[Result "1/2-1/2"]

1. e4 a6 2. e5 d5 3. exd6 Nf6 4. dxc7 Nc6 ; rest of line is a comment
5. cxd8=N g6 6. Nf3 Bg7 7. Bc4 O-O 8. O-O Rxd8 1/2-1/2

[Event "black wins"]
[Result "0-1"]

1. f3 e5 2. g4 Qh4# 0-1

[Event "unfinished"]
[Result "*"]

1. e4 *
"""


def replay_san_on_oracle(san_moves, resolve_san):
    """plays SAN tokens on the oracle board through the product's SAN resolver; returns (RefGame, [uci moves])"""
    g = ref_path.RefGame(chess960=False, start_id=None)
    ucis = []
    for san in san_moves:
        b = g.board
        legal = list(b.legal_moves)

        def piece_at(sq, b=b):
            p = b.piece_at(sq)
            return (p.piece_type, p.color) if p else None
        k = resolve_san(san, [(m.from_square, m.to_square, m.promotion) for m in legal], piece_at)
        ucis.append(legal[k].uci())
        g.move_piece(legal[k])
    return g, ucis
