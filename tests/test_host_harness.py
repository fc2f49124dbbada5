"""The DEVICE chess / tree code (sigma-zero_b200/csrc/chess.cuh, tree.cuh) compiled for the host and checked
against the oracle.  No GPU needed; this is the same source the CUDA kernels call."""
import ctypes

import numpy as np
import pytest

from oracle import hash_eval, numerics, ref_path
import chess
from tests import util


@pytest.fixture(scope="module")
def H():
    return util.build_host_harness()


def _from_board(H, b):
    bb, turn, rw, rb, ep, hm, ply = b._export()
    return H.hh_set((ctypes.c_uint64 * 12)(*bb), turn, rw, rb, ep, hm, ply, int(b.chess960))


@pytest.mark.parametrize("fen,c960,expected", util.PERFT_KATS)
def test_perft_kats(H, fen, c960, expected):
    g = _from_board(H, chess.Board(fen, chess960=c960))
    depth = 4 if expected[3] < 5_000_000 else 3
    assert [H.hh_perft(g, d + 1) for d in range(depth)] == expected[:depth]
    H.hh_free(g)


def test_chess960_starts_match_oracle(H):
    for sid in range(0, 960, 7):
        g = H.hh_new(sid)
        b = chess.Board.from_chess960_pos(sid)
        assert [H.hh_perft(g, d) for d in (1, 2, 3)] == [b.perft(d) for d in (1, 2, 3)], sid
        H.hh_free(g)


def test_random_playouts_match_oracle(H):
    """legal index sets, packed planes (8-ply history, repetition and castling planes), outcomes, codec round trip"""
    rng = np.random.default_rng(11)
    positions = 0
    for gi in range(16):
        c960 = gi % 2 == 1
        sid = int(rng.integers(960)) if c960 else -1
        g = H.hh_new(sid)
        og = ref_path.RefGame(chess960=c960, start_id=sid if c960 else None)
        while True:
            b = og.board
            legal = list(b.legal_moves)
            buf = (ctypes.c_uint16 * 256)()
            n = H.hh_legal(g, buf)
            assert list(buf[:n]) == util.legal_indices(og)
            pl = (ctypes.c_uint64 * 119)()
            H.hh_planes(g, pl)
            assert np.array_equal(np.array(pl[:], dtype=np.uint64), hash_eval.pack_planes(og.get_representation()))
            assert H.hh_outcome(g) == util.outcome_code(b)
            assert H.hh_codec_roundtrip(g) == 0
            positions += 1
            if b.outcome() is not None or len(b.move_stack) > 400:
                break
            m = legal[rng.integers(len(legal))]
            assert H.hh_push_index(g, ref_path.move_to_index(m, b.turn)) == 0
            og.move_piece(m)
        H.hh_free(g)
    assert positions > 1000


def test_illegal_move_rejected(H):
    g = H.hh_new(-1)
    assert H.hh_push_index(g, ref_path.move_to_index(chess.Move.from_uci("e2e5"), True)) < 0
    assert H.hh_push_index(g, ref_path.move_to_index(chess.Move.from_uci("e2e4"), True)) == 0
    H.hh_free(g)


def test_puct_cascade_noise_bit_exact(H, golden_dir):
    import os
    z = np.load(os.path.join(golden_dir, "numerics.npz"))
    for x, y in zip(z["sum_x"], z["sum_y"]):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert H.hh_cascade_sum(x.ctypes.data_as(ctypes.c_void_p)) == y
    # the sparse walk over the legal-move bitset (what k_finish runs) == the dense cascade on the masked vector == torch.sum
    rng = np.random.default_rng(11)
    for trial in range(300):
        n_legal = int(rng.integers(1, 219)) if trial % 7 else 0
        idx = np.sort(rng.choice(4672, size=n_legal, replace=False))
        if trial % 5 == 0 and n_legal:                      # clustered like real move lists, tail block included
            idx = np.unique(np.clip(idx % 300 + 4672 - 300 * (trial % 2) - (0 if trial % 2 else 4372), 0, 4671))
        x = np.zeros(4672, dtype=np.float32)
        vals = rng.random(len(idx)).astype(np.float32) ** 4
        vals[rng.random(len(idx)) < 0.05] = 0.0             # zero priors at legal moves
        x[idx] = vals
        bits = np.zeros(73 * 64, dtype=np.uint8)
        bits[idx] = 1
        mask = np.packbits(bits, bitorder="little").view("<u8").copy()
        dense = H.hh_cascade_sum(x.ctypes.data_as(ctypes.c_void_p))
        sparse = H.hh_cascade_sum_sparse(x.ctypes.data_as(ctypes.c_void_p), mask.ctypes.data_as(ctypes.c_void_p))
        assert np.float32(dense) == np.float32(sparse) == numerics.cascade_sum(x), trial
    for i in range(len(z["puct_len"])):
        for k in range(int(z["puct_len"][i])):
            got = H.hh_puct(int(z["puct_n"][i][k]), float(z["puct_w"][i][k]), float(z["puct_p"][i][k]),
                            int(z["puct_N"][i]), float(z["puct_C"][i]))
            assert np.float32(got) == z["puct_ucb"][i][k]
    for p in np.random.default_rng(0).random(200).astype(np.float32):
        assert np.float32(H.hh_noisy_prior(float(p))) == numerics.noisy_prior(np.array([p]))[0]


def test_hash_evaluator_twin(H):
    rng = np.random.default_rng(3)
    for _ in range(5):
        words = rng.integers(0, 2 ** 63, size=119, dtype=np.uint64)
        pol = np.zeros(4672, dtype=np.float32)
        val = ctypes.c_float()
        H.hh_hash_eval(words.ctypes.data_as(ctypes.c_void_p), pol.ctypes.data_as(ctypes.c_void_p), ctypes.byref(val))
        p, v = hash_eval.evaluate_words(words)
        assert np.array_equal(pol, p) and np.float32(val.value) == v
