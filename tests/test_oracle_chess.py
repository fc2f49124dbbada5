"""Known-answer tests that pin the oracle's rules engine (oracle/chess_oracle.c + oracle/chess).

The reference ships no tests, and python-chess is not installable here, so the pins are external:
public perft tables (standard + Chess960), Scharnagl anchors, and rule-by-rule cases for the
python-chess semantics the hot path touches.
"""
import copy

import pytest

from oracle import ref_path  # noqa: F401  (puts oracle/ on sys.path so `chess` resolves to the stand-in)
import chess

PERFT = [
    ("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", False, [20, 400, 8902, 197281]),
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", False, [48, 2039, 97862]),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", False, [14, 191, 2812, 43238, 674624]),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", False, [6, 264, 9467, 422333]),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", False, [44, 1486, 62379]),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", False, [46, 2079, 89890]),
    ("bqnb1rkr/pp3ppp/3ppn2/2p5/5P2/P2P4/NPP1P1PP/BQ1BNRKR w HFhf - 2 9", True, [21, 528, 12189, 326672]),
    ("2nnrbkr/p1qppppp/8/1ppb4/6PP/3PP3/PPP2P2/BQNNRBKR w HEhe - 1 9", True, [21, 807, 18002, 667366]),
    ("b1q1rrkb/pppppppp/3nn3/8/P7/1PPP4/4PPPP/BQNNRKRB w GE - 1 9", True, [20, 479, 10471, 273318]),
]


@pytest.mark.parametrize("fen,c960,expected", PERFT)
def test_perft_tables(fen, c960, expected):
    b = chess.Board(fen, chess960=c960)
    assert [b.perft(d + 1) for d in range(len(expected))] == expected


def test_perft_start_depth5():
    assert chess.Board().perft(5) == 4865609


def test_scharnagl_anchors():
    assert chess.chess960_backrank(0) == "BBQNNRKR"
    assert chess.chess960_backrank(518) == "RNBQKBNR"
    assert chess.chess960_backrank(959) == "RKRNNQBB"
    seen = set()
    for i in range(960):
        row = chess.chess960_backrank(i)
        assert sorted(row) == sorted("RNBQKBNR")
        b1, b2 = [k for k, c in enumerate(row) if c == "B"]
        assert (b1 + b2) % 2 == 1                          # opposite-coloured bishops
        assert row.index("R") < row.index("K") < row.rindex("R")
        seen.add(row)
    assert len(seen) == 960
    assert chess.Board.from_chess960_pos(518).perft(3) == 8902


def test_castling_notation_by_mode():
    v = chess.Board("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1")
    assert {"e1g1", "e1c1"} <= {m.uci() for m in v.legal_moves}
    x = chess.Board("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1", chess960=True)
    ucis = {m.uci() for m in x.legal_moves}
    assert {"e1h1", "e1a1"} <= ucis and "e1g1" not in ucis
    v.push(chess.Move.from_uci("e1g1"))
    assert str(v).splitlines()[-1] == "R . . . . R K ."
    assert not v.has_kingside_castling_rights(chess.WHITE) and v.has_queenside_castling_rights(chess.BLACK)
    assert v.move_stack[-1].uci() == "e1g1"


def test_chess960_castling_first_move():
    # king f1, rook g1 style starts allow castling on move one in some ids; just check agreement of rights
    for sid in (0, 100, 518, 700, 959):
        b = chess.Board.from_chess960_pos(sid)
        assert b.has_kingside_castling_rights(chess.WHITE) and b.has_queenside_castling_rights(chess.BLACK)
        assert b.chess960


def test_castling_through_attack_and_960_backrank_discovery():
    # rook on f8 attacks f1: white may not castle king-side, may castle queen-side
    b = chess.Board("5r2/8/8/8/8/8/8/R3K2R w KQ - 0 1")
    ucis = {m.uci() for m in b.legal_moves}
    assert "e1g1" not in ucis and "e1c1" in ucis
    # Chess960: king b1, rook a1, enemy rook/queen on the back rank behind the rook's start is harmless,
    # but a queen on e1..h1 line attacking c1 after the rook leaves matters; construct: K on b1, R on a1,
    # black rook on h1 with d1..g1 empty -> after a-side castling king c1, rook d1 blocks h1: legal.
    b = chess.Board("4k3/8/8/8/8/8/8/RK5r w A - 0 1", chess960=True)
    assert b.is_check()                                     # h1 rook checks b1: no castling out of check
    assert "b1a1" not in {m.uci() for m in b.legal_moves}


def test_en_passant_rules():
    b = chess.Board()
    for u in ("e2e4", "a7a6", "e4e5", "d7d5"):
        b.push(chess.Move.from_uci(u))
    assert b.ep_square == chess.parse_square("d6")
    assert "e5d6" in {m.uci() for m in b.legal_moves}
    # the classic rank skewer: capturing e.p. would expose the king on the fifth rank
    b = chess.Board("8/8/8/K2pP2r/8/8/8/7k w - d6 0 1")
    assert "e5d6" not in {m.uci() for m in b.legal_moves}
    # ep square is set even when no capture is possible (python-chess keeps it after every double push)
    b = chess.Board()
    b.push(chess.Move.from_uci("e2e4"))
    assert b.ep_square == chess.parse_square("e3")


def test_repetition_semantics():
    b = chess.Board()
    seq = ["g1f3", "g8f6", "f3g1", "f6g8"]
    for u in seq:
        b.push(chess.Move.from_uci(u))
    assert b.is_repetition(2) and not b.is_repetition(3)
    for u in seq:
        b.push(chess.Move.from_uci(u))
    assert b.is_repetition(3) and not b.is_repetition(5)
    assert not b.is_game_over()
    for u in seq + seq:
        b.push(chess.Move.from_uci(u))
    assert b.is_repetition(5) and b.is_game_over() and b.result() == "1/2-1/2"
    # an irreversible move (pawn push) stops the walk
    b = chess.Board()
    for u in ["g1f3", "g8f6", "f3g1", "f6g8", "e2e4", "e7e5"]:
        b.push(chess.Move.from_uci(u))
    assert not b.is_repetition(2)
    # losing castling rights makes the otherwise identical position different
    b = chess.Board("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1")
    for u in ["h1g1", "h8g8", "g1h1", "g8h8"]:
        b.push(chess.Move.from_uci(u))
    assert not b.is_repetition(2)
    for u in ["h1g1", "h8g8", "g1h1", "g8h8"]:
        b.push(chess.Move.from_uci(u))
    assert b.is_repetition(2)


def test_game_over_rules():
    assert chess.Board("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1").result() == "1/2-1/2"          # stalemate
    m = chess.Board("7k/6Q1/6K1/8/8/8/8/8 b - - 0 1")
    assert m.is_game_over() and m.outcome().winner is chess.WHITE and m.result() == "1-0"
    assert chess.Board("8/8/8/4k3/8/8/8/4K3 w - - 0 1").is_game_over()                  # K v K
    assert chess.Board("8/8/8/4k3/8/8/5N2/4K3 w - - 0 1").is_game_over()                # K+N v K
    assert not chess.Board("8/8/8/4k3/8/8/4NN2/4K3 w - - 0 1").is_game_over()           # K+N+N v K
    assert chess.Board("8/8/8/4kb2/8/8/5B2/4K3 w - - 0 1").is_game_over() is False      # opposite-coloured bishops
    assert chess.Board("8/8/8/4k1b1/8/8/5B2/4K3 w - - 0 1").is_game_over()              # same-coloured bishops
    assert not chess.Board("8/8/8/4k3/8/8/4P3/4K3 w - - 0 1").is_game_over()
    assert not chess.Board("8/8/8/4k3/8/8/8/R3K3 w - - 149 1").is_game_over()
    assert chess.Board("8/8/8/4k3/8/8/8/R3K3 w - - 150 1").is_game_over()               # 75-move rule
    assert not chess.Board("8/8/8/4k3/8/8/8/R3K3 w - - 100 1").is_game_over()           # 50-move is only claimable


def test_deepcopy_keeps_move_stack():
    b = chess.Board()
    for u in ["g1f3", "g8f6", "f3g1"]:
        b.push(chess.Move.from_uci(u))
    c = copy.deepcopy(b)
    c.push(chess.Move.from_uci("f6g8"))
    assert c.is_repetition(2) and len(b.move_stack) == 3 and len(c.move_stack) == 4
    assert [m.uci() for m in c.move_stack] == ["g1f3", "g8f6", "f3g1", "f6g8"]
