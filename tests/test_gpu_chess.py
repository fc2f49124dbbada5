"""GPU parity: the CUDA chess engine (through the C ABI) against the oracle -- bit-exact legal move sets
(perft), encoded planes, move indices, repetition / castling planes and game outcomes (config c1)."""
import os

import numpy as np
import pytest

from oracle import hash_eval, ref_path
import chess
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from sigma_zero_b200.engine import Engine
    e = Engine(max_games=1024, max_searches=64)
    yield e
    e.close()


@pytest.mark.parametrize("fen,c960,expected", util.PERFT_KATS)
def test_perft_published_tables(eng, fen, c960, expected):
    from sigma_zero_b200 import _lib
    pos = util.wire_pos(chess.Board(fen, chess960=c960), _lib)
    assert [eng.perft(pos, d + 1) for d in range(len(expected))] == expected


def test_perft_all_960_starts_vs_oracle(eng):
    from sigma_zero_b200 import _lib
    for sid in range(960):
        b = chess.Board.from_chess960_pos(sid)
        pos = util.wire_pos(b, _lib)
        depth = 4 if sid % 8 == 0 else 3
        assert [eng.perft(pos, d) for d in range(1, depth + 1)] == [b.perft(d) for d in range(1, depth + 1)], sid
    # id 518 in chess960 mode reproduces the standard numbers; deeper on a few ids
    for sid in (518, 0, 959, 341):
        b = chess.Board.from_chess960_pos(sid)
        assert eng.perft(util.wire_pos(b, _lib), 5) == b.perft(5), sid
    assert eng.perft(util.wire_pos(chess.Board.from_chess960_pos(518), _lib), 5) == 4865609


def test_lockstep_playouts_vs_oracle(eng):
    """96 games (vanilla + Chess960) played in lock step with random moves; every ply compares legal index
    lists, packed planes, legal masks and outcomes of all games."""
    rng = np.random.default_rng(2)
    G = 96
    specs = [(g % 2 == 1, int(rng.integers(960)) if g % 2 else 518, []) for g in range(G)]
    games = util.setup_games(eng, specs)
    alive = list(range(G))
    checked = 0
    for ply in range(220):
        idx, cnt = eng.legal_moves()
        planes, mask = eng.encode()
        pos = eng.positions()
        who, mv = [], []
        for g in range(G):
            og = games[g]
            exp = util.legal_indices(og)
            assert list(idx[g, :cnt[g]]) == exp, (g, ply, og.board.fen())
            assert np.array_equal(planes[g], hash_eval.pack_planes(og.get_representation())), (g, ply)
            from sigma_zero_b200.engine import child_indices
            assert list(child_indices(mask[g])) == exp
            assert pos[g].outcome == util.outcome_code(og.board), (g, ply, og.board.fen())
            assert pos[g].ply == len(og.board.move_stack) and pos[g].halfmove_clock == og.board.halfmove_clock
            assert pos[g].rep_flags == int(og.board.is_repetition(2)) + 2 * int(og.board.is_repetition(3)) or ply == 0
            checked += 1
            if g in alive and og.board.outcome() is None:
                legal = list(og.board.legal_moves)
                m = legal[rng.integers(len(legal))]
                who.append(g)
                mv.append(ref_path.move_to_index(m, og.board.turn))
                og.move_piece(m)
            elif g in alive:
                alive.remove(g)
        if not who:
            break
        eng.push(who, mv)
    assert checked > 5000


def test_codec_golden(eng, golden_dir):
    z = np.load(os.path.join(golden_dir, "codec.npz"))
    n = len(z["sid"])
    off = np.concatenate([[0], np.cumsum(z["idx_len"])])
    for lo in range(0, n, 512):
        hi = min(n, lo + 512)
        specs = [(bool(z["c960"][i]), int(z["sid"][i]), str(z["moves"][i]).split()) for i in range(lo, hi)]
        util.setup_games(eng, specs)
        idx, cnt = eng.legal_moves()
        planes, _ = eng.encode()
        pos = eng.positions()
        for k, i in enumerate(range(lo, hi)):
            assert np.array_equal(idx[k, :cnt[k]].astype(np.int64), z["idx_flat"][off[i]:off[i + 1]].astype(np.int64)), i
            assert np.array_equal(planes[k], z["planes"][i]), i
            assert (pos[k].outcome != 0) == bool(z["terminal"][i])
            assert (-1 if pos[k].outcome == 1 else 0) == int(z["value"][i])


def test_illegal_move_is_rejected(eng):
    eng.reset([-1, -1])
    bad = ref_path.move_to_index(chess.Move.from_uci("e2e5"), True)
    good = ref_path.move_to_index(chess.Move.from_uci("e2e4"), True)
    with pytest.raises(ValueError, match="Invalid move"):
        eng.push([0, 1], [good, bad])
    pos = eng.positions()
    assert pos[0].ply == 1 and pos[1].ply == 0          # the legal one went through, the illegal one left its game alone


def test_unpack_planes_matches_reference_float_layout(eng):
    import ctypes
    import torch
    eng.reset([-1, 77])
    eng.push([0, 1], [ref_path.move_to_index(chess.Move.from_uci("e2e4"), True), util.legal_indices(util.oracle_game(True, 77))[0]])
    planes, _ = eng.encode()
    d_pl = torch.from_numpy(planes.astype(np.int64)).cuda()
    out = torch.empty((2, 119, 8, 8), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rc = eng.lib.szb_unpack_planes_f32(eng._h, 2, ctypes.c_void_p(d_pl.data_ptr()), ctypes.c_void_p(out.data_ptr()))
    assert rc == 0
    eng.synchronize()
    og = util.oracle_game(False, 518, ["e2e4"])
    assert np.array_equal(out[0].cpu().numpy(), og.get_representation().astype(np.float32))
