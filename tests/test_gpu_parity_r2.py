"""GPU parity at the sizes BASELINE.json states (round-2 additions; everything goes through the C ABI):

  c1  perft 1-5 of all 960 Chess960 start positions + the vanilla one vs the oracle (fixture from oracle/make_perft960.py + live sample)
  c2  bf16 tcgen05 network vs torch fp32 at 1024 boards through the chunked 512-board launches (logits AND a peaked policy);
      fp32 "identical network outputs" visit-count parity at 800 simulations on 8 positions; bf16-vs-fp32 visit distributions of an
      800-simulation search at 1024 concurrent games
  c3  4096 concurrent Chess960 games x 800 simulations: size-independent properties
  and the step restructuring of round 2: one tree kernel + one tower launch with fused heads must build exactly the trees of the
  separate phase kernels.

Measured error figures are written to gpurun_out/parity_r2.json (copied to profiles/ by hand)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import hash_eval, ref_path
import chess
from tests import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(key, value):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, "parity_r2.json")
    cur = json.load(open(p)) if os.path.exists(p) else {}
    cur[key] = value
    json.dump(cur, open(p, "w"), indent=1)


def _random_specs(n, seed, max_plies=40):
    """(c960, start id, uci moves) of n mid-game positions, vanilla and Chess960 alternating, none finished"""
    rng = np.random.default_rng(seed)
    specs = []
    for g in range(n):
        c960 = g % 2 == 1
        sid = int(rng.integers(960)) if c960 else 518
        og = util.oracle_game(c960, sid)
        moves = []
        for _ in range(int(rng.integers(0, max_plies))):
            legal = list(og.board.legal_moves)
            if not legal or og.board.outcome() is not None:
                break
            m = legal[rng.integers(len(legal))]
            moves.append(m.uci())
            og.move_piece(m)
        while og.board.outcome() is not None and moves:
            moves = moves[:-2]
            og = util.oracle_game(c960, sid, moves)
        specs.append((c960, sid, moves))
    return specs


# ---------------------------------------------------------------------------------------------------------------------------
# c1
# ---------------------------------------------------------------------------------------------------------------------------
def test_perft_depth_1_to_5_all_960_starts_vs_oracle(golden_dir):
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.engine import Engine
    table = np.load(os.path.join(golden_dir, "perft960.npz"))["nodes"]
    eng = Engine(max_games=4, max_searches=4)
    for sid in list(range(960)) + [-1]:
        b = chess.Board() if sid < 0 else chess.Board.from_chess960_pos(sid)
        pos = util.wire_pos(b, _lib)
        got = [eng.perft(pos, d) for d in range(1, 6)]
        assert got == table[sid if sid >= 0 else 960].tolist(), (sid, got)
    # and against the oracle itself, live, on a few ids (the fixture is the oracle's output, committed)
    for sid in (7, 518, 902):
        b = chess.Board.from_chess960_pos(sid)
        assert eng.perft(util.wire_pos(b, _lib), 5) == b.perft(5), sid
    eng.close()


# ---------------------------------------------------------------------------------------------------------------------------
# the one-kernel tree step + fused heads vs the separate phase kernels
# ---------------------------------------------------------------------------------------------------------------------------
def _trees(eng, games):
    out = []
    for g in games:
        t = eng.tree_export(g)
        out.append({k: (v.copy() if hasattr(v, "copy") else v) for k, v in t.items()})
    return out


@pytest.mark.parametrize("evaluator_name", ["hash", "bf16", "fp32"])
def test_fused_step_equals_phase_kernels(evaluator_name):
    """normal searches run k_tree_step (+ for the bf16 network one tower launch whose last epilogue emits priors and values);
    with profiling on, the library runs select / expand / planes / tower / value head / softmax / finish as separate kernels.
    Same trees, node for node: visit counts, value sums (fp64), priors (fp32 bits), child sets"""
    from sigma_zero_b200.engine import EVAL_HASH, EVAL_NET_BF16, EVAL_NET_FP32, Engine
    ev = {"hash": EVAL_HASH, "bf16": EVAL_NET_BF16, "fp32": EVAL_NET_FP32}[evaluator_name]
    n, S = (37, 120) if evaluator_name != "fp32" else (6, 40)
    specs = _random_specs(n, seed=31)
    # a mate-in-one position: terminal leaves inside the tree (no evaluation, constant value)
    eng = Engine(max_games=n + 1, max_searches=S, cohorts=1)
    if evaluator_name != "hash":
        torch.manual_seed(0)
        eng.load_state_dict(ref_path.build_policy_nn().eval().state_dict())
    util.setup_games(eng, specs + [(False, 518, ["f2f3", "e7e5", "g2g4"])])
    res = {}
    for prof in (False, True):
        eng.set_profiling(prof)
        v, c, val = eng.search(S, 2.0, True, ev, want_value=True)
        res[prof] = (v, c, val, _trees(eng, range(0, n + 1, 6)))
    eng.set_profiling(False)
    eng.close()
    assert np.array_equal(res[False][0], res[True][0]) and np.array_equal(res[False][1], res[True][1])
    assert np.array_equal(res[False][2], res[True][2])
    for ta, tb in zip(res[False][3], res[True][3]):
        for k in ta:
            assert np.array_equal(np.asarray(ta[k]), np.asarray(tb[k])), k
    assert (res[False][0].sum(axis=1) == S - 1).all()


def test_fused_heads_equal_softmax_of_logits():
    """the priors the fused epilogue writes for the legal moves are bit-equal to k_softmax over the full logits row: a search on
    the bf16 network must give the visit counts of the restated reference search fed szb_net_forward's policy / value"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine, child_indices
    torch.manual_seed(1)
    sd = ref_path.build_policy_nn().eval().state_dict()
    specs = _random_specs(3, seed=5, max_plies=30)
    eng = Engine(max_games=4, max_searches=64, cohorts=1)
    eng.load_state_dict(sd)
    aux = Engine(max_games=2, max_searches=4)
    aux.load_state_dict(sd)

    def ev(planes):
        pol, val = aux.net_forward(hash_eval.pack_planes(planes)[None], EVAL_NET_BF16)
        return pol[0], val[0]

    games = util.setup_games(eng, specs)
    for learning in (False, True):
        visits, child, _ = eng.search(64, 2.0, learning, EVAL_NET_BF16)
        for g, og in enumerate(games):
            _, root = ref_path.search(og, 64, 2.0, ev, learning=learning)
            idx = [c.index for c in root.children]
            assert list(child_indices(child[g])) == idx, (g, learning)
            assert visits[g, idx].tolist() == [c.n for c in root.children], (g, learning)
    eng.close()
    aux.close()


# ---------------------------------------------------------------------------------------------------------------------------
# c2: network at size
# ---------------------------------------------------------------------------------------------------------------------------
def test_bf16_vs_torch_fp32_at_1024_boards_chunked_path():
    """1024 boards = two 512-board tower launches over the same activation rows.  Logits are compared relative to the largest
    |logit| (<= 3e-2; measured 0.9e-2).  The post-softmax policy is compared twice: on the network as initialised (north star:
    <= 2e-2 abs -- but every prior is ~2e-4 there, so the error is ALSO bounded relative to the largest prior), and on a PEAKED
    head (conv_p2 scaled ~25x until the mean largest prior exceeds 0.3, logits up to +-31).  For the peaked head the bound is what
    the logit error allows, |dp| <= p (1 - p) |dlogit| <= 0.25 x the largest logit error, and 5e-2 abs: a bf16 tower cannot hold
    2e-2 on a head scaled that far (measured 3.3e-2 max, 8e-6 mean) -- recorded in profiles/, not hidden"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine
    torch.manual_seed(7)
    model = ref_path.build_policy_nn().eval()
    g = torch.Generator().manual_seed(3)
    for m in model.modules():                                    # non-trivial BatchNorm statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
            m.running_mean = 0.1 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var = 0.5 + torch.rand(m.running_var.shape, generator=g)
    specs = _random_specs(128, seed=77, max_plies=60)
    planes = np.stack([util.oracle_game(c, s, mv).get_representation() for c, s, mv in specs])
    x = torch.from_numpy(planes.astype(np.float32))
    with torch.no_grad():
        rl, rv = model(x, inference=False)
    peak_scale = 1.0
    for _ in range(40):                                          # scale the policy head until the policy is peaked
        if float(torch.softmax(rl * peak_scale, dim=1).max(dim=1).values.mean()) > 0.3:
            break
        peak_scale *= 1.5
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd["conv_p2.weight"] = sd["conv_p2.weight"] * peak_scale
    sd["conv_p2.bias"] = sd["conv_p2.bias"] * peak_scale
    rl = (rl * peak_scale).numpy()
    rp = torch.softmax(torch.from_numpy(rl), dim=1).numpy()
    rv = rv.numpy().ravel()
    order = np.random.default_rng(1).permutation(1024) % 128     # every position 8 times, at rows of both chunks
    packed = np.stack([hash_eval.pack_planes(p) for p in planes])[order]
    eng = Engine(max_games=1024, max_searches=4)
    eng.load_state_dict(sd)
    l16, v16 = eng.net_forward(packed, EVAL_NET_BF16, logits=True)
    p16, _ = eng.net_forward(packed, EVAL_NET_BF16)
    eng.load_state_dict(model.state_dict())                      # the head as initialised (flat policy)
    p16_flat, _ = eng.net_forward(packed, EVAL_NET_BF16)
    eng.close()
    rp_flat = torch.softmax(torch.from_numpy(rl / peak_scale), dim=1).numpy()[order]
    err_flat = np.abs(p16_flat - rp_flat)
    scale = float(np.abs(rl).max())
    err_l = np.abs(l16 - rl[order])
    err_p = np.abs(p16 - rp[order])
    err_v = np.abs(v16 - rv[order])
    stats = {"boards": 1024, "peak_scale": peak_scale, "max_abs_logit": scale, "logit_err_max": float(err_l.max()),
             "logit_err_mean": float(err_l.mean()), "logit_err_max_rel": float(err_l.max() / scale),
             "mean_max_prior": float(rp.max(axis=1).mean()), "policy_err_max": float(err_p.max()), "policy_err_mean": float(err_p.mean()),
             "value_err_max": float(err_v.max()), "value_err_mean": float(err_v.mean()),
             "argmax_agree": float((p16.argmax(1) == rp[order].argmax(1)).mean()),
             "flat_head_policy_err_max": float(err_flat.max()), "flat_head_max_prior": float(rp_flat.max()),
             "flat_head_policy_err_max_rel_to_max_prior": float((err_flat.max(axis=1) / rp_flat.max(axis=1)).max())}
    print("bf16 vs torch fp32 at 1024 boards:", stats)
    _record("bf16_vs_torch_fp32_1024_boards", stats)
    assert stats["mean_max_prior"] > 0.3
    assert stats["logit_err_max_rel"] <= 3e-2
    assert stats["value_err_max"] <= 2e-2
    assert stats["flat_head_policy_err_max"] <= 2e-2 and stats["flat_head_policy_err_max_rel_to_max_prior"] <= 5e-2
    assert stats["policy_err_max"] <= min(5e-2, 0.25 * stats["logit_err_max"] * 1.05 + 1e-3) and stats["policy_err_mean"] <= 1e-4
    # copies of one position in different rows / chunks agree bit for bit (batch invariance through the chunked path)
    first = {}
    for row, k in enumerate(order):
        if k in first:
            assert np.array_equal(l16[row], l16[first[k]]) and v16[row] == v16[first[k]]
        else:
            first[k] = row


def test_fp32_identical_outputs_visit_parity_800_sims_8_positions():
    """north star: "given identical network outputs in fp32 mode, MCTS visit counts and selected moves are bit-exact" -- at the
    reference's evaluation budget (800 simulations, C = 2), on 8 mid-game positions, 4 of them Chess960, with the REAL network
    (fp32 parity path) on the GPU side and the restated reference search on the other, fed the GPU's fp32 forward of the planes
    the oracle itself encodes (mcts.py:39-122)"""
    from sigma_zero_b200.engine import EVAL_NET_FP32, Engine, child_indices
    torch.manual_seed(5)
    sd = ref_path.build_policy_nn().eval().state_dict()
    specs = _random_specs(8, seed=404, max_plies=50)
    eng = Engine(max_games=8, max_searches=800, cohorts=1)
    eng.load_state_dict(sd)
    aux = Engine(max_games=2, max_searches=4)
    aux.load_state_dict(sd)

    def ev(planes):
        pol, val = aux.net_forward(hash_eval.pack_planes(planes)[None], EVAL_NET_FP32)
        return pol[0], val[0]

    games = util.setup_games(eng, specs)
    visits, child, _ = eng.search(800, 2.0, True, EVAL_NET_FP32)
    for g, og in enumerate(games):
        probs, root = ref_path.search(og, 800, 2.0, ev, learning=True)
        idx = [c.index for c in root.children]
        assert list(child_indices(child[g])) == idx, g
        assert visits[g, idx].tolist() == [c.n for c in root.children], g
        assert int(np.argmax(visits[g])) == idx[int(np.argmax([c.n for c in root.children]))]       # selected move (first maximum)
    eng.close()
    aux.close()


def test_c2_bf16_vs_fp32_visit_distributions_at_size():
    """config c2: 1024 concurrent vanilla games x 800 simulations on the bf16 tcgen05 network (two cohorts, 512-board launches),
    "visit counts checked vs fp32 reference": the fp32 parity network searches the first 64 of the same games (a game's search does
    not depend on the batch it runs in -- test_cohort_split / test_batch_invariance -- and 819,200 fp32 SIMT evaluations would take
    minutes).  Reported: total-variation distance per game and agreement of the selected move; thresholds = observed + margin"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, EVAL_NET_FP32, Engine
    torch.manual_seed(0)
    sd = ref_path.build_policy_nn().eval().state_dict()
    G, S, SUB = 1024, 800, 64
    eng = Engine(max_games=G, max_searches=S)
    eng.load_state_dict(sd)
    eng.reset([-1] * G)
    rng = np.random.default_rng(12)
    plies = rng.integers(0, 16, G)
    for ply in range(int(plies.max())):
        idx, cnt = eng.legal_moves()
        who = [g for g in range(G) if ply < plies[g] and cnt[g] > 0]
        eng.push(who, [int(idx[g, (5 * g + 3 * ply) % cnt[g]]) for g in who])
    pos = eng.positions()
    running = np.array([p.outcome == 0 for p in pos])              # a random prefix can end a game (fool's mate)
    v16, c16, _ = eng.search(S, 2.0, False, EVAL_NET_BF16)
    assert (v16.sum(axis=1)[running] == S - 1).all() and (v16.sum(axis=1)[~running] == 0).all() and running.sum() >= G - 8
    eng.close()
    sub = Engine(max_games=SUB, max_searches=S, cohorts=1)
    sub.load_state_dict(sd)
    sub.set_positions([pos[g] for g in range(SUB)])            # same positions, empty history: search both sides from those
    v16s, _, _ = sub.search(S, 2.0, False, EVAL_NET_BF16)
    v32, c32, _ = sub.search(S, 2.0, False, EVAL_NET_FP32)
    sub.close()
    keep = running[:SUB]
    v16s, v32 = v16s[keep], v32[keep]
    tv = 0.5 * np.abs(v16s.astype(np.float64) - v32).sum(1) / (S - 1.0)
    same_top = float((v16s.argmax(1) == v32.argmax(1)).mean())
    # top move of one search among the other's top 3
    top3 = float(np.mean([v16s[g].argmax() in np.argsort(-v32[g].astype(np.int64), kind="stable")[:3] for g in range(len(v32))]))
    # games that have not moved carry no history planes, so the copy searched alone must equal the one inside the big batch
    kept = np.nonzero(keep)[0]
    fresh = [k for k, g in enumerate(kept) if plies[g] == 0]
    assert fresh and all(np.array_equal(v16[kept[k]], v16s[k]) for k in fresh)
    stats = {"games_bf16": G, "games_fp32": SUB, "sims": S, "tv_mean": float(tv.mean()), "tv_max": float(tv.max()),
             "tv_median": float(np.median(tv)), "same_top_move": same_top, "bf16_top_in_fp32_top3": top3}
    print("c2 bf16 vs fp32 visit distributions:", stats)
    _record("c2_bf16_vs_fp32_visits", stats)
    assert tv.mean() < 0.15 and same_top >= 0.6


# ---------------------------------------------------------------------------------------------------------------------------
# c3: at size
# ---------------------------------------------------------------------------------------------------------------------------
def test_c3_full_size_properties():
    """config c3: 4096 concurrent Chess960 games x 800 simulations on one GPU (two cohorts, eight 512-board tower launches per step).
    What must hold at any size: child visits of every running game sum to num_searches - 1, root children = legal moves, a repeated
    search reproduces the counts bit for bit, games searched inside the batch equal the same games searched alone, counters add up"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine, child_indices
    torch.manual_seed(0)
    sd = ref_path.build_policy_nn().eval().state_dict()
    G, S = 4096, 800
    eng = Engine(max_games=G, max_searches=S)
    eng.load_state_dict(sd)
    ids = [(37 * g + 11) % 960 for g in range(G)]
    eng.reset(ids)
    for ply in range(6):
        idx, cnt = eng.legal_moves()
        who = [g for g in range(G) if (g + ply) % 3 and cnt[g] > 0]
        eng.push(who, [int(idx[g, (11 * g + 5 * ply) % cnt[g]]) for g in who])
    s0 = eng.stats()
    v, c, _ = eng.search(S, 2.0, True, EVAL_NET_BF16)
    s1 = eng.stats()
    assert (v.sum(axis=1) == S - 1).all()
    assert s1["simulations"] - s0["simulations"] == G * S
    assert s1["evaluations"] - s0["evaluations"] + s1["terminal_visits"] - s0["terminal_visits"] == G * S
    idx, cnt = eng.legal_moves()
    for g in range(0, G, 97):
        assert list(child_indices(c[g])) == idx[g, :cnt[g]].tolist(), g
    pos = eng.positions()
    hist = eng.encode(want_mask=False)[0]
    v2, c2, _ = eng.search(S, 2.0, True, EVAL_NET_BF16)
    assert np.array_equal(v, v2) and np.array_equal(c, c2)
    eng.close()
    # a handful of the games alone (those that have not moved yet carry no history, so set_positions reproduces them exactly)
    alone = [g for g in range(0, G, 251) if pos[g].ply == 0][:8] or [0]
    small = Engine(max_games=len(alone), max_searches=S, cohorts=1)
    small.load_state_dict(sd)
    small.reset([ids[g] for g in alone])
    va, ca, _ = small.search(S, 2.0, True, EVAL_NET_BF16)
    small.close()
    for k, g in enumerate(alone):
        if pos[g].ply == 0:
            assert np.array_equal(va[k], v[g]) and np.array_equal(ca[k], c[g]), g
    assert hist.shape == (G, 119)


# ---------------------------------------------------------------------------------------------------------------------------
# boundary behaviour added in round 2
# ---------------------------------------------------------------------------------------------------------------------------
def test_create_rejects_arena_beyond_32_bit_edge_indices():
    from sigma_zero_b200.engine import Engine, SzbError
    with pytest.raises(SzbError, match="2\\^31"):
        Engine(max_games=65536, max_searches=800, edges_per_node=48)


def test_bad_checkpoint_keeps_the_loaded_network():
    """szb_net_load validates the whole state_dict before it tears the working network down"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine, SzbError
    torch.manual_seed(0)
    sd = ref_path.build_policy_nn().eval().state_dict()
    eng = Engine(max_games=4, max_searches=4)
    eng.load_state_dict(sd)
    planes = hash_eval.pack_planes(util.oracle_game(False, -1).get_representation())[None]
    before = eng.net_forward(planes, EVAL_NET_BF16)
    bad = dict(sd)
    del bad["fc_v2.bias"]
    with pytest.raises(SzbError, match="fc_v2.bias"):
        eng.load_state_dict(bad)
    after = eng.net_forward(planes, EVAL_NET_BF16)
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    eng.close()


# ---------------------------------------------------------------------------------------------------------------------------
# SURVEY 8f rows 2 and 4 against the ORACLE (round 1 compared the engine with itself)
# ---------------------------------------------------------------------------------------------------------------------------
def _gpu_fp32_evaluator(state_dict):
    from sigma_zero_b200.engine import EVAL_NET_FP32, Engine
    eng = Engine(max_games=2, max_searches=4)
    eng.load_state_dict(state_dict)

    def ev(planes):
        pol, val = eng.net_forward(hash_eval.pack_planes(planes)[None], EVAL_NET_FP32)
        return pol[0], val[0]

    return eng, ev


def test_arena_moves_equal_reference_search_argmax():
    """arena.play_match (test_update.py:26-83: learning=False, arg-max of the visit counts, first maximum) against the restated
    reference search: every move of both networks in both games must be the oracle's first arg-max, given identical network
    outputs (fp32 parity networks)"""
    from sigma_zero_b200.arena import play_match
    from sigma_zero_b200.network import policyNN
    torch.manual_seed(1)
    a = policyNN({"precision": "fp32"}).eval()
    torch.manual_seed(2)
    b = policyNN({"precision": "fp32"}).eval()
    args = {"C": 2, "num_searches": 40}
    out = play_match(a, b, 2, args, c960=False, max_plies=4)
    assert out["plies"] == 4 and all(len(m) == 4 for m in out["moves"])
    ea, eva = _gpu_fp32_evaluator(a.state_dict())
    eb, evb = _gpu_fp32_evaluator(b.state_dict())
    for g in range(2):
        og = util.oracle_game(False, -1)
        for ply, idx in enumerate(out["moves"][g]):
            a_to_move = out["a_is_white"][g] == (ply % 2 == 0)
            _, root = ref_path.search(og, 40, 2, eva if a_to_move else evb, learning=False)
            best = root.children[int(np.argmax([c.n for c in root.children]))]
            assert best.index == idx, (g, ply)
            og.move_piece(best.move)
    ea.close()
    eb.close()


def test_playtensor_best_move_equals_reference_first_max():
    """play.py:40-43 get_best_move = max(probs, key=probs.get) over MCTS0.search: same visit fractions and the same selected move
    as the restated reference search with identical network outputs"""
    from sigma_zero_b200.play import PlayTensor
    torch.manual_seed(4)
    p = PlayTensor(num_searches=48, precision="fp32")
    eng, ev = _gpu_fp32_evaluator(p.network.state_dict())
    played = []
    for u in ("d2d4", None, "c2c4", None):
        if u is None:
            og = util.oracle_game(False, -1, played)
            probs = p.mcts.search(p.board, verbose=False)
            ref_probs, root = ref_path.search(og, 48, 2, ev, learning=False)
            assert [m.uci() for m in probs] == [m.uci() for m in ref_probs] and list(probs.values()) == list(ref_probs.values())
            best = p.get_best_move()
            assert best.uci() == max(ref_probs, key=ref_probs.get).uci()
            p.move(best)
            played.append(best.uci())
        else:
            p.move(u)
            played.append(u)
    eng.close()


# ---------------------------------------------------------------------------------------------------------------------------
# the cluster-resident tower (batches of at most 18 boards) against the CTA-pair tower
# ---------------------------------------------------------------------------------------------------------------------------
def test_cluster_tower_bit_identical(monkeypatch):
    """k_tower_cl (one cluster of 8 or 16 CTAs per board, M = 64 MMAs, activations exchanged through distributed shared memory, heads inside
    the launch) runs the same MMAs per output in the same K order as k_tower_tc2: logits, softmax policy and value must agree bit
    for bit, for 1, 5 and 18 boards, with non-trivial BatchNorm statistics"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine
    torch.manual_seed(3)
    model = ref_path.build_policy_nn().eval()
    g = torch.Generator().manual_seed(13)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
            m.running_mean = 0.1 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var = 0.5 + torch.rand(m.running_var.shape, generator=g)
    specs = _random_specs(18, seed=23, max_plies=50)
    packed = np.stack([hash_eval.pack_planes(util.oracle_game(c, s, mv).get_representation()) for c, s, mv in specs])
    for n in (1, 5, 18):
        outs = []
        for cluster, size in (("0", "0"), ("18", "8"), ("18", "16"), ("18", "0")):     # pair kernel | clusters of 8 | of 16 | automatic
            monkeypatch.setenv("SZB_TOWER_CLUSTER", cluster)
            monkeypatch.setenv("SZB_TOWER_CLUSTER_SIZE", size)
            eng = Engine(max_games=n, max_searches=4)
            eng.load_state_dict(model.state_dict())
            outs.append(eng.net_forward(packed[:n], EVAL_NET_BF16, logits=True) + eng.net_forward(packed[:n], EVAL_NET_BF16))
            eng.close()
        for other in outs[1:]:
            for a, b in zip(outs[0], other):
                assert np.array_equal(a, b), n
        assert np.abs(outs[1][0]).max() > 0.1 and np.isfinite(outs[1][0]).all()


def test_cluster_tower_search_equals_pair_tower(monkeypatch):
    """searches of a few games (fused two-launch steps) with the cluster-resident tower and with the CTA-pair tower: same visit
    counts, same root values; and the profiled (phase-kernel) form agrees as well"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine
    torch.manual_seed(0)
    sd = ref_path.build_policy_nn().eval().state_dict()
    specs = _random_specs(7, seed=61) + [(False, 518, ["f2f3", "e7e5", "g2g4"])]
    res = []
    for cluster, prof in (("0", False), ("18", False), ("18", True)):
        monkeypatch.setenv("SZB_TOWER_CLUSTER", cluster)
        eng = Engine(max_games=8, max_searches=150, cohorts=1)
        eng.load_state_dict(sd)
        util.setup_games(eng, specs)
        eng.set_profiling(prof)
        res.append(eng.search(150, 2.0, True, EVAL_NET_BF16, want_value=True))
        eng.set_profiling(False)
        eng.close()
    for r in res[1:]:
        for a, b in zip(res[0], r):
            assert np.array_equal(a, b)
    assert (res[0][0].sum(axis=1) == 149).all()


# ---------------------------------------------------------------------------------------------------------------------------
# device-resident weight path (train_RL.py:211-227: the weights reach the workers)
# ---------------------------------------------------------------------------------------------------------------------------
def test_device_weight_load_equals_host_load_and_refills_in_place():
    """szb_net_load_device (flat fp32 buffer already on the GPU, BatchNorm folded and packed by kernels) must give exactly the network
    szb_net_load (host tensors) gives: equal digests of the packed weights, bit-equal fp32 and bf16 outputs; a second, different
    set of weights refills the same context in place and a reload of the first restores the first digest"""
    from sigma_zero_b200 import runtime
    from sigma_zero_b200.engine import EVAL_NET_BF16, EVAL_NET_FP32, Engine
    models = []
    for seed in (0, 1):
        torch.manual_seed(seed)
        m = ref_path.build_policy_nn().eval()
        g = torch.Generator().manual_seed(40 + seed)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.data = 0.5 + torch.rand(mod.weight.shape, generator=g)
                mod.bias.data = 0.2 * torch.randn(mod.bias.shape, generator=g)
                mod.running_mean = 0.1 * torch.randn(mod.running_mean.shape, generator=g)
                mod.running_var = 0.5 + torch.rand(mod.running_var.shape, generator=g)
        models.append(m)
    planes = np.stack([hash_eval.pack_planes(util.oracle_game(c, s, mv).get_representation()) for c, s, mv in _random_specs(5, seed=2)])
    host = Engine(max_games=8, max_searches=4)
    host.load_state_dict(models[0].state_dict())
    dev = Engine(max_games=8, max_searches=4)
    dev.load_flat_device(*runtime.flat_weights(models[0], torch.device("cuda", 0)))
    d0 = dev.net_checksum()
    assert d0 == host.net_checksum()
    for ev in (EVAL_NET_FP32, EVAL_NET_BF16):
        a, b = host.net_forward(planes, ev, logits=True), dev.net_forward(planes, ev, logits=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    dev.load_flat_device(*runtime.flat_weights(models[1], torch.device("cuda", 0)))
    assert dev.net_checksum() != d0
    dev.load_flat_device(*runtime.flat_weights(models[0], torch.device("cuda", 0)))
    assert dev.net_checksum() == d0
    # the fold itself against torch: fp32 path within the north-star tolerance with non-trivial BatchNorm statistics
    x = torch.from_numpy(np.stack([util.oracle_game(c, s, mv).get_representation() for c, s, mv in _random_specs(5, seed=2)]).astype(np.float32))
    with torch.no_grad():
        rp, rv = models[0](x, inference=True)
    p, v = dev.net_forward(planes, EVAL_NET_FP32)
    assert np.abs(p - rp.numpy()).max() <= 1e-5 and np.abs(v - rv.numpy().ravel()).max() <= 1e-5
    host.close()
    dev.close()
