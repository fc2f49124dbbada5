"""GPU parity: batched MCTS (tree kernels + in-tree move generation + plane encoding) against the reference
search.  With the integer-hash evaluator both sides see bit-identical "network" outputs, so visit counts and
child sets must agree exactly (mcts.py:39-122 semantics incl. the zero-prior drop and the constant noise)."""
import os

import numpy as np
import pytest

from oracle import hash_eval, ref_path
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from sigma_zero_b200.engine import Engine
    e = Engine(max_games=256, max_searches=800)
    yield e
    e.close()


def _check(eng, games, n, C, learning, visits, child, tag=""):
    from sigma_zero_b200.engine import child_indices
    for g, og in enumerate(games):
        _, root = ref_path.search(og, n, C, hash_eval.evaluator, learning=learning)
        exp_idx = [c.index for c in root.children]
        exp_n = [c.n for c in root.children]
        got_idx = list(child_indices(child[g]))
        assert got_idx == exp_idx, (tag, g)
        assert visits[g, exp_idx].tolist() == exp_n, (tag, g)
        assert int(visits[g].sum()) == sum(exp_n)


def test_search_golden_visit_counts(eng, golden_dir):
    """the 50 cases recorded from the UNMODIFIED reference MCTS0.search (tests/golden/search.npz)"""
    from sigma_zero_b200.engine import EVAL_HASH, child_indices
    z = np.load(os.path.join(golden_dir, "search.npz"))
    off = np.concatenate([[0], np.cumsum(z["idx_len"])])
    groups = {}
    for i in range(len(z["sid"])):
        groups.setdefault((int(z["n"][i]), float(z["C"][i]), bool(z["learning"][i])), []).append(i)
    for (n, C, learning), members in groups.items():
        specs = [(bool(z["c960"][i]), int(z["sid"][i]), str(z["moves"][i]).split()) for i in members]
        util.setup_games(eng, specs)
        visits, child, _ = eng.search(n, C, learning, EVAL_HASH)
        for k, i in enumerate(members):
            exp_idx = z["idx_flat"][off[i]:off[i + 1]].astype(np.int64)
            assert np.array_equal(child_indices(child[k]), exp_idx), (i, n, C, learning)
            assert np.array_equal(visits[k, exp_idx], z["visits_flat"][off[i]:off[i + 1]].astype(np.uint32)), i


@pytest.mark.parametrize("learning", [False, True])
def test_search_800_sims_vs_live_oracle(eng, learning):
    """config-c2-shaped search (800 simulations, C=2) on a handful of games, vanilla and Chess960, mid-game"""
    from sigma_zero_b200.engine import EVAL_HASH
    rng = np.random.default_rng(9 + learning)
    specs = []
    for g in range(4):
        c960 = g % 2 == 1
        sid = int(rng.integers(960)) if c960 else 518
        og = util.oracle_game(c960, sid)
        moves = []
        for _ in range(int(rng.integers(0, 40))):
            legal = list(og.board.legal_moves)
            if not legal or og.board.outcome() is not None:
                break
            m = legal[rng.integers(len(legal))]
            moves.append(m.uci())
            og.move_piece(m)
        if og.board.outcome() is not None:
            moves = moves[:-2]
        specs.append((c960, sid, moves))
    games = util.setup_games(eng, specs)
    visits, child, _ = eng.search(800, 2.0, learning, EVAL_HASH)
    _check(eng, games, 800, 2, learning, visits, child, "800")
    st = eng.stats()
    assert st["simulations"] >= 800 * len(specs)


def test_selfplay_teacher_forced(eng):
    """self-play plies on the GPU (search -> sample -> push); at every ply the oracle, fed the move the GPU
    chose, must reproduce the position, and a fresh oracle search must reproduce the visit counts"""
    from sigma_zero_b200.engine import EVAL_HASH
    specs = [(False, 518, []), (True, 3, []), (True, 944, []), (False, 518, ["g2g4", "e7e5", "f2f3"])]
    games = util.setup_games(eng, specs)
    n = 48
    for ply in range(12):
        visits, child, _ = eng.search(n, 2.0, True, EVAL_HASH)
        _check(eng, games, n, 2, True, visits, child, "ply%d" % ply)
        moves, active = eng.selfplay_ply(n, 2.0, True, EVAL_HASH, seed=1234, sample=True)
        for g, og in enumerate(games):
            if og.board.outcome() is not None:
                assert moves[g] == -1
                continue
            assert visits[g, moves[g]] > 0                  # sampled proportionally to visits => never an unvisited child
            qp = {(m.from_square, m.to_square) for m in og.board.legal_moves if m.promotion == 5}
            og.move_piece(ref_path.index_to_move(int(moves[g]), og.board.turn, qp))
        planes, _ = eng.encode()
        for g, og in enumerate(games):
            assert np.array_equal(planes[g], hash_eval.pack_planes(og.get_representation()))


def test_terminal_root_and_single_search(eng):
    """search on a finished game returns no children (mcts.py:113-122 with an empty child list); num_searches=1
    leaves every child at zero visits"""
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.engine import EVAL_HASH
    import chess
    mate = chess.Board("7k/6Q1/6K1/8/8/8/8/8 b - - 0 1")
    start = chess.Board()
    eng.set_positions([util.wire_pos(mate, _lib), util.wire_pos(start, _lib)])
    visits, child, _ = eng.search(10, 2.0, False, EVAL_HASH)
    assert visits[0].sum() == 0 and child[0].sum() == 0
    assert visits[1].sum() == 9
    visits, child, _ = eng.search(1, 2.0, False, EVAL_HASH)
    assert visits[1].sum() == 0 and child[1].sum() != 0


def test_finished_games_take_no_slot(eng):
    """games that are over take no slot of the search batch (no step kernels, no network rows): the running games of a mixed
    batch get exactly the visit counts of a batch holding only them, a finished root still ends with visit_count 1 + num_searches
    and value_sum num_searches * (-1 | 0) (mcts.py:46,104-109), and the counters say which simulations were evaluations"""
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.engine import EVAL_HASH
    import chess
    mate = chess.Board("7k/6Q1/6K1/8/8/8/8/8 b - - 0 1")
    stale = chess.Board("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1")
    start = chess.Board()
    kiwi = chess.Board("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
    S = 40
    eng.set_positions([util.wire_pos(b, _lib) for b in (start, kiwi)])
    v_ref, c_ref, _ = eng.search(S, 2.0, True, EVAL_HASH)
    s0 = eng.stats()
    eng.set_positions([util.wire_pos(b, _lib) for b in (mate, start, stale, kiwi, mate)])
    v, c, rv = eng.search(S, 2.0, True, EVAL_HASH, want_value=True)
    s1 = eng.stats()
    assert np.array_equal(v[[1, 3]], v_ref) and np.array_equal(c[[1, 3]], c_ref)
    assert v[[0, 2, 4]].sum() == 0 and c[[0, 2, 4]].sum() == 0
    assert rv[0] == -1.0 and rv[2] == 0.0
    assert s1["simulations"] - s0["simulations"] == 5 * S
    assert s1["terminal_visits"] - s0["terminal_visits"] >= 3 * S
    assert s1["evaluations"] - s0["evaluations"] <= 2 * S
    for g, (n, w) in {0: (1 + S, -float(S)), 2: (1 + S, 0.0), 4: (1 + S, -float(S))}.items():
        t = eng.tree_export(g)
        assert t["root_visits"] == n and t["root_value_sum"] == w and len(t["node_first"]) == 1
    # all games over: nothing to search, still well defined
    eng.set_positions([util.wire_pos(mate, _lib), util.wire_pos(stale, _lib)])
    v, c, _ = eng.search(S, 2.0, False, EVAL_HASH)
    assert v.sum() == 0 and c.sum() == 0
    moves, active = eng.selfplay_ply(S, 2.0, True, EVAL_HASH, seed=1, sample=True)
    assert active == 0 and (moves == -1).all()


def test_cohort_split_does_not_change_results():
    """the two-cohort pipeline (two game halves stepping on two streams) must give exactly the visit counts of the single
    batch -- with the hash evaluator (exact) and with the bf16 network (batch-invariant kernels)"""
    import torch
    from sigma_zero_b200.engine import EVAL_HASH, EVAL_NET_BF16, Engine
    torch.manual_seed(0)
    sd = ref_path.build_policy_nn().eval().state_dict()
    n = 30                                                  # split 16 / 14: ragged second cohort, 4-board tile boundary inside
    out = {}
    for c in (1, 2):
        e = Engine(max_games=n, max_searches=40, cohorts=c)
        e.load_state_dict(sd)
        e.reset([(-1 if g % 3 else (53 * g) % 960) for g in range(n)])
        for ply in range(3):                                # a few plies so that positions differ between games
            idx, cnt = e.legal_moves()
            e.push(list(range(n)), [int(idx[g, (7 * g + ply) % cnt[g]]) for g in range(n)])
        vh, ch, _ = e.search(40, 2.0, True, EVAL_HASH)
        vn, cn, _ = e.search(40, 2.0, True, EVAL_NET_BF16)
        out[c] = (vh, ch, vn, cn)
        e.close()
    for a, b in zip(out[1], out[2]):
        assert np.array_equal(a, b)


def test_full_size_c2_properties():
    """BASELINE config c2 at full size (1024 concurrent games x 800 simulations, bf16 tcgen05 network, two cohorts, 512-board tower
    launches): what must hold whatever the size -- every running game's child visits sum to num_searches - 1 (mcts.py:46,113-122),
    the root children are exactly the legal moves (no softmax underflow with this network), a second identical search and a
    single-cohort engine reproduce the visit counts bit for bit, and the counters add up"""
    import torch
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine, child_indices
    torch.manual_seed(0)
    sd = ref_path.build_policy_nn().eval().state_dict()
    G, S = 1024, 800
    rng = np.random.default_rng(5)
    out = []
    for cohorts in (0, 1):
        e = Engine(max_games=G, max_searches=S, cohorts=cohorts)
        e.load_state_dict(sd)
        e.reset([-1] * G)
        plies = rng.integers(0, 12, G) if cohorts == 0 else plies
        for ply in range(int(plies.max())):                  # a few random legal plies so that the games differ
            idx, cnt = e.legal_moves()
            who = [g for g in range(G) if ply < plies[g] and cnt[g] > 0]
            e.push(who, [int(idx[g, (7 * g + 3 * ply) % cnt[g]]) for g in who])
        s0 = e.stats()
        v, c, _ = e.search(S, 2.0, True, EVAL_NET_BF16)
        s1 = e.stats()
        out.append((v, c))
        if cohorts == 0:
            v2, c2, _ = e.search(S, 2.0, True, EVAL_NET_BF16)
            assert np.array_equal(v, v2) and np.array_equal(c, c2)
            idx, cnt = e.legal_moves()
            for g in range(0, G, 37):
                assert list(child_indices(c[g])) == idx[g, :cnt[g]].tolist(), g
        assert (v.sum(axis=1) == S - 1).all()
        assert s1["simulations"] - s0["simulations"] == G * S
        assert s1["evaluations"] - s0["evaluations"] + s1["terminal_visits"] - s0["terminal_visits"] == G * S
        e.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_multi_leaf_virtual_loss_mode():
    """opt-in szb_config.leaves_per_tree = K > 1 (K simulations of a tree in flight per step, virtual loss): NOT the reference's
    algorithm, so no visit-count parity -- what must hold: exactly num_searches simulations per game (child visits sum to
    num_searches - 1, root visit count 1 + num_searches), the same root children as the exact search, bit-identical repeats,
    a consistent tree (every expanded node's edge count = 1 + its children's, no pending marker or virtual loss left behind),
    terminal nodes and finished games handled, counters that add up"""
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.engine import EVAL_HASH, Engine
    import chess
    S = 200
    boards = [chess.Board(), chess.Board("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"),
              chess.Board("6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1"),             # mate in one: terminal nodes inside the tree
              chess.Board("7k/6Q1/6K1/8/8/8/8/8 b - - 0 1"),                    # finished game
              chess.Board("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1")]
    pos = [util.wire_pos(b, _lib) for b in boards]
    ref = Engine(max_games=8, max_searches=S)
    ref.set_positions(pos)
    v1, c1, _ = ref.search(S, 2.0, True, EVAL_HASH)
    ref.close()
    live = [0, 1, 2, 4]
    for K in (2, 5, 8):
        e = Engine(max_games=8, max_searches=S, leaves_per_tree=K)
        e.set_positions(pos)
        s0 = e.stats()
        v, c, _ = e.search(S, 2.0, True, EVAL_HASH)
        s1 = e.stats()
        assert (v[live].sum(axis=1) == S - 1).all(), (K, v.sum(axis=1))
        assert v[3].sum() == 0 and np.array_equal(c, c1)
        assert s1["simulations"] - s0["simulations"] == len(boards) * S
        assert (s1["evaluations"] - s0["evaluations"]) + (s1["terminal_visits"] - s0["terminal_visits"]) == len(boards) * S
        for g in live:
            t = e.tree_export(g)
            assert t["root_visits"] == 1 + S
            n_nodes = len(t["node_first"])
            kids = t["edge_child"]
            assert ((kids == -1) | ((kids > 0) & (kids < n_nodes))).all()          # no pending marker left
            assert sorted(kids[kids > 0].tolist()) == list(range(1, n_nodes))      # every node hangs on exactly one edge
            assert (t["edge_visits"] >= 0).all()
            for i in range(1, n_nodes):
                lo, cnt = t["node_first"][i], t["node_count"][i]
                n_edge = t["edge_visits"][t["node_parent_edge"][i]]
                if cnt:
                    assert n_edge == 1 + t["edge_visits"][lo:lo + cnt].sum(), (K, g, i)
                else:
                    assert t["node_terminal"][i] and n_edge >= 1
            unvisited = kids == -1
            assert (t["edge_visits"][unvisited] == 0).all() and (t["edge_value_sum"][unvisited] == 0).all()
        v2, c2, _ = e.search(S, 2.0, True, EVAL_HASH)
        assert np.array_equal(v, v2) and np.array_equal(c, c2)
        tv = 0.5 * np.abs(v[live].astype(np.float64) - v1[live]).sum(axis=1) / (S - 1)
        print("leaves_per_tree=%d: total-variation distance to the exact search %s" % (K, np.round(tv, 3)))
        assert tv.max() < 0.6
        moves, active = e.selfplay_ply(S, 2.0, True, EVAL_HASH, seed=3, sample=True)
        assert active >= 3 and moves[3] == -1 and (moves[live] >= 0).all()
        e.close()
    # the mate in one is found: Rd8# (d1d8) collects the most visits in every mode
    best = int(np.argmax(v[2]))
    assert best == int(np.argmax(v1[2]))
