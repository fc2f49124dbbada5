"""BASELINE config c1 without a GPU: tests/golden/perft960.npz (perft 1-5 of all 960 Chess960 start positions and the vanilla one,
counted by the oracle's C engine, oracle/make_perft960.py) against the live oracle on a sample, and the HOST build of the device
rules code (sigma-zero_b200/csrc/chess.cuh) against the fixture on every start position."""
import os

import numpy as np
import pytest

from oracle import ref_path  # noqa: F401
import chess
from tests import util


@pytest.fixture(scope="module")
def table(golden_dir):
    z = np.load(os.path.join(golden_dir, "perft960.npz"))
    assert z["ids"].tolist() == list(range(960)) + [-1] and z["nodes"].shape == (961, 5)
    return z["nodes"]


def test_fixture_matches_live_oracle_on_a_sample(table):
    assert table[518].tolist() == [20, 400, 8902, 197281, 4865609]          # published numbers (id 518 = the standard array)
    assert table[960].tolist() == table[518].tolist()                       # vanilla castling rules, same counts
    rng = np.random.default_rng(0)
    for sid in [0, 959] + [int(x) for x in rng.integers(0, 960, 6)]:
        b = chess.Board.from_chess960_pos(sid)
        assert [b.perft(d) for d in range(1, 6)] == table[sid].tolist(), sid


def test_host_build_of_device_rules_vs_fixture_all_960(table):
    """every start position to depth 3, every 8th to depth 4 through the g++ build of chess.cuh (tests/host_harness)"""
    H = util.build_host_harness()
    for sid in list(range(960)) + [-1]:
        g = H.hh_new(sid)
        depth = 4 if sid % 8 == 0 else 3
        assert [H.hh_perft(g, d) for d in range(1, depth + 1)] == table[sid if sid >= 0 else 960, :depth].tolist(), sid
        H.hh_free(g)
