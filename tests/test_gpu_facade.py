"""GPU parity through the drop-in facade (the reference's own call surface): MCTS0.search / Node / ChessTensor /
policyNN / sim.play_game running on libszb200, against the restated reference search fed with THE SAME network
outputs (the GPU's fp32 forward) -- the north-star's "given identical network outputs in fp32 mode, MCTS visit
counts and selected moves are bit-exact"."""
import numpy as np
import pytest
import torch

from oracle import hash_eval, ref_path
from tests import util

pytestmark = pytest.mark.gpu

OPENING = ["e2e4", "c7c5", "g1f3", "d7d6", "d2d4", "c5d4", "f3d4", "g8f6", "b1c3", "a7a6"]


@pytest.fixture(scope="module")
def fp32_model():
    from sigma_zero_b200.network import policyNN
    torch.manual_seed(5)
    return policyNN({"precision": "fp32"}).eval()


@pytest.fixture(scope="module")
def evaluator(fp32_model):
    """oracle-side evaluator: the GPU's own fp32 forward for the planes the oracle produces (identical network outputs)"""
    from sigma_zero_b200.engine import EVAL_NET_FP32, Engine
    eng = Engine(max_games=2, max_searches=4)
    eng.load_state_dict(fp32_model.state_dict())

    def ev(planes):
        pol, val = eng.net_forward(hash_eval.pack_planes(planes)[None], EVAL_NET_FP32)
        return pol[0], val[0]

    yield ev
    eng.close()


@pytest.mark.parametrize("learning", [False, True])
def test_mcts0_search_matches_reference_with_identical_network_outputs(fp32_model, evaluator, learning):
    from sigma_zero_b200 import chess_compat as cc
    from sigma_zero_b200.chess_tensor import ChessTensor
    from sigma_zero_b200.mcts import MCTS0
    args = {"C": 2, "num_searches": 60}
    game = ChessTensor()
    for u in OPENING:
        game.move_piece(cc.Move.from_uci(u))
    mcts = MCTS0(game, args, fp32_model)
    probs = mcts.search(game.board, verbose=False, learning=learning)
    og = util.oracle_game(False, -1, OPENING)
    ref_probs, ref_root = ref_path.search(og, args["num_searches"], args["C"], evaluator, learning=learning)
    assert [m.uci() for m in probs] == [m.uci() for m in ref_probs]            # same children, ascending move index
    assert list(probs.values()) == list(ref_probs.values())                      # same visit fractions, exactly
    assert max(probs, key=probs.get).uci() == max(ref_probs, key=ref_probs.get).uci()   # same selected move
    # Node view of the device tree (mctsnode.py fields)
    root = mcts.tree()
    assert root.visit_count == 1 + args["num_searches"] and root.visit_count == ref_root.n
    assert [c.visit_count for c in root.children] == [c.n for c in ref_root.children]
    assert [c.value_sum for c in root.children] == [c.w for c in ref_root.children]
    assert [np.float32(c.prior) for c in root.children] == [np.float32(c.prior) for c in ref_root.children]
    assert root.select().action_taken.uci() == ref_path._select(ref_root, args["C"]).move.uci()
    with pytest.raises(RuntimeError):
        root.expand([])


def test_chess_tensor_facade_matches_oracle_game():
    from sigma_zero_b200 import chess_compat as cc
    from sigma_zero_b200.chess_tensor import ChessTensor
    game = ChessTensor()
    og = util.oracle_game(False, -1)
    for u in OPENING + ["c1g5", "e7e6", "d1d2", "f8e7", "e1c1"]:         # includes castling (vanilla: e1c1)
        assert sorted(m.uci() for m in game.get_moves()) == sorted(m.uci() for m in og.board.legal_moves)
        game.move_piece(cc.Move.from_uci(u))
        og.move_piece(util.chess.Move.from_uci(u))
        assert np.array_equal(game.get_representation().numpy(), og.get_representation())
        assert game.get_value_and_terminated() == og.value_and_terminated()
        assert game.board.turn == og.board.turn and game.board.halfmove_clock == og.board.halfmove_clock
        for colour in (True, False):
            assert game.board.has_kingside_castling_rights(colour) == og.board.has_kingside_castling_rights(colour)
            assert game.board.has_queenside_castling_rights(colour) == og.board.has_queenside_castling_rights(colour)
    # the absolute-orientation stacks the reference keeps (chess_tensor.py:123-129): un-flip the player's view
    view = og.get_representation()
    if og.board.turn:
        assert np.array_equal(game.representation.numpy(), view[:, ::-1, :])
    else:
        assert np.array_equal(game.black_representation.numpy(), view[:, :, ::-1])
    with pytest.raises(ValueError, match="Invalid move"):
        game.move_piece(cc.Move.from_uci("a1a8"))


def test_policynn_facade_forward_matches_torch(fp32_model):
    ref = ref_path.build_policy_nn().eval()
    ref.load_state_dict(fp32_model.state_dict())                 # same 252-key layout (network.py / play.py:25-28)
    x = torch.from_numpy(np.stack([util.oracle_game(False, -1, OPENING[:k]).get_representation() for k in (0, 3, 10)])).float()
    p, v = fp32_model(x, inference=True)
    with torch.no_grad():
        rp, rv = ref(x, inference=True)
    assert p.shape == (3, 4672) and v.shape == (3, 1)
    assert (p - rp).abs().max().item() <= 1e-5 and (v - rv).abs().max().item() <= 1e-5


def test_play_game_facade_schema():
    """sim.play_game drop-in (bf16 network, 8 searches/move): the reference's history schema, consistent rewards"""
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.sim import generate_training_data, selfplay_batch
    torch.manual_seed(0)
    model = policyNN({}).eval()
    args = {"C": 2, "num_searches": 8, "num_selfPlay_iterations": 6, "chess960": True}
    games, counters = selfplay_batch(model, args, 6, c960=True, seed=3, max_plies=12)
    assert len(games) == 6 and counters["plies"] == 12
    for h in games:
        n = len(h["actions"])
        assert n == len(h["states"]) == len(h["colours"]) == len(h["rewards"]) == 12
        assert h["states"][0].shape == (119, 8, 8) and h["states"][0].dtype == torch.bool
        assert h["colours"][0] is True and h["colours"][1] is False
        for probs in h["actions"]:
            assert abs(sum(probs.values()) - 1.0) < 1e-9
        assert h["result"] == "*" and set(h["rewards"]) == {0}
    merged = {}
    out = generate_training_data(model, num_games=2, args={"C": 2, "num_searches": 4}, return_dict=merged, c960=False)
    assert len(merged) == 1 and set(out) == {"states", "actions", "rewards", "colours"}
    assert len(out["states"]) == len(out["actions"]) == len(out["rewards"]) == len(out["colours"]) > 0


def test_bf16_visit_counts_track_the_fp32_search():
    """config c2's check: the bf16 tcgen05 network drives the same search as the fp32 parity network up to rounding.
    Visit distributions are compared per game (total-variation distance) -- exact equality is not expected."""
    from sigma_zero_b200.engine import EVAL_NET_BF16, EVAL_NET_FP32, Engine
    torch.manual_seed(0)
    model = ref_path.build_policy_nn().eval()
    eng = Engine(max_games=32, max_searches=200)
    eng.load_state_dict(model.state_dict())
    specs = [(g % 2 == 1, 518 if g % 2 == 0 else (37 * g) % 960, OPENING[: (g % 6)] if g % 2 == 0 else []) for g in range(32)]
    util.setup_games(eng, specs)
    v16, _, _ = eng.search(200, 2.0, False, EVAL_NET_BF16)
    v32, _, _ = eng.search(200, 2.0, False, EVAL_NET_FP32)
    assert (v16.sum(1) == 199).all() and (v32.sum(1) == 199).all()
    tv = 0.5 * np.abs(v16.astype(np.float64) - v32).sum(1) / 199.0
    same_top = float((v16.argmax(1) == v32.argmax(1)).mean())
    print("bf16 vs fp32 search: mean TV %.4f  max TV %.4f  same top move %.2f" % (tv.mean(), tv.max(), same_top))
    import json, os
    out = os.path.join(util.ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump({"games": 32, "sims": 200, "tv_mean": float(tv.mean()), "tv_max": float(tv.max()), "same_top_move": same_top},
              open(os.path.join(out, "parity_bf16_vs_fp32_32x200.json"), "w"))
    # observed on B200 (profiles/r02b_parity.json has the at-size figures): mean TV 1.6e-4 at 1024 x 800; the bounds leave a margin
    # for other weights / positions but no longer admit an unrelated search
    assert tv.mean() < 0.03 and tv.max() < 0.25 and same_top >= 0.8
    eng.close()


def test_sharded_selfplay_plays_the_same_games():
    """a job split over ranks (train_RL.shard_of blocks, game_id_base = first global id) must produce, game for
    game, what one process produces -- start positions, sampled moves, visit distributions.  Two 'ranks' are emulated one
    after the other on this GPU."""
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.sim import selfplay_batch
    from sigma_zero_b200.train_RL import shard_of
    torch.manual_seed(0)
    model = policyNN({}).eval()
    args = {"C": 2, "num_searches": 12}
    whole, _ = selfplay_batch(model, args, 10, c960=True, seed=7, max_plies=6)
    parts = []
    for rank in range(2):
        lo, hi = shard_of(10, rank, 2)
        games, _ = selfplay_batch(model, args, hi - lo, c960=True, seed=7, max_plies=6, game_id_base=lo)
        parts += games
    assert len(parts) == len(whole) == 10
    for a, b in zip(whole, parts):
        assert len(a["actions"]) == len(b["actions"]) == 6
        for pa, pb in zip(a["actions"], b["actions"]):
            assert [m.uci() for m in pa] == [m.uci() for m in pb] and list(pa.values()) == list(pb.values())
        for sa, sb in zip(a["states"], b["states"]):
            assert torch.equal(sa, sb)


def test_arena_match_between_two_networks():
    """SURVEY 8f rank 2 (test_update.py): two networks, every game concurrently, arg-max moves, learning=False.  The moves of
    the first ply are checked against a stand-alone search of the same positions with the same weights."""
    from sigma_zero_b200.arena import play_match, update_model
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine
    from sigma_zero_b200.network import policyNN
    torch.manual_seed(1)
    a = policyNN({}).eval()
    torch.manual_seed(2)
    b = policyNN({}).eval()
    args = {"C": 2, "num_searches": 24}
    out = play_match(a, b, 6, args, c960=True, seed=3, max_plies=5)
    assert out["plies"] == 5 and len(out["results"]) == 6 and out["a_is_white"] == [True, False] * 3
    assert all(r == "*" for r in out["results"]) and out["score_a"] == 0.0
    again = play_match(a, b, 6, args, c960=True, seed=3, max_plies=5)
    assert again == out                                             # deterministic
    assert update_model(b, a, matches=1, args=args, max_plies=4) is False      # nothing finished in 4 plies -> no promotion
    # first ply of a vanilla game: network a is White in game 0, network b in game 1 -- each move must be the first arg-max
    # of a stand-alone search of the start position with that network's weights
    one = play_match(a, b, 2, args, c960=False, max_plies=1)
    assert one["plies"] == 1
    for g, model in enumerate((a, b)):
        eng = Engine(max_games=1, max_searches=24, cohorts=1)
        eng.load_state_dict(model.state_dict())
        eng.reset([-1])
        eng.search(24, 2.0, False, EVAL_NET_BF16, want_visits=False, want_children=False)
        idx, vis, cnt = eng.root_children()
        assert one["moves"][g] == [int(idx[0, int(np.argmax(vis[0, :cnt[0]]))])]
        eng.close()


def test_playtensor_single_game_api():
    """one game, model replies with the arg-max move of its search"""
    from sigma_zero_b200.play import PlayTensor
    torch.manual_seed(4)
    p = PlayTensor(num_searches=16)
    assert p.check_if_end() is None
    p.move("e2e4")
    reply = p.model_move()
    assert p.board.turn is True and len(p.board.move_stack) == 2 and reply.uci() == p.board.move_stack[-1].uci()
    with pytest.raises(ValueError, match="Invalid move"):
        p.move("e2e4")
    p.start_new_game()
    assert len(p.board.move_stack) == 0


def test_selfplay_train_selfplay_loop():
    """the reference's outer loop in miniature (train_RL.py:205-264): self-play on the CUDA engine -> packed records -> the
    reference's fine-tuning step in torch on the GPU -> the engine picks the new weights up for the next self-play"""
    from sigma_zero_b200 import records
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.sim import selfplay_batch
    from sigma_zero_b200.train_RL import train_on_records
    torch.manual_seed(0)
    model = policyNN({}).eval()
    args = {"C": 2, "num_searches": 16}
    games, _ = selfplay_batch(model, args, 8, c960=False, seed=1, max_plies=4)
    for g in games:                                         # unfinished games carry zero rewards: give the value head something to fit
        g["rewards"] = [1 if i % 2 == 0 else -1 for i in range(len(g["rewards"]))]
    rec = records.pack_records(games)
    assert rec["states"].shape == (32, 119)
    x = torch.from_numpy(records.unpack_states(rec, [0, 5])).float()
    before_p, before_v = model(x, inference=True)            # CUDA kernels, old weights
    hist = train_on_records(model, rec, epochs=3, batch_size=16, device="cuda")
    assert len(hist) == 6 and all(np.isfinite(h).all() for h in hist) and sum(hist[-1]) < sum(hist[0])
    after_p, after_v = model(x, inference=True)              # CUDA kernels again: weights re-sent because they changed
    assert not torch.equal(before_p, after_p) and not torch.equal(before_v, after_v)
    with torch.no_grad():
        ref_p, ref_v = model.forward_torch(x.cuda(), inference=True)        # same weights through torch (eval-mode BN)
    assert (after_p - ref_p.cpu()).abs().max().item() <= 2e-2 and (after_v - ref_v.cpu()).abs().max().item() <= 2e-2
    games2, _ = selfplay_batch(model, args, 8, c960=False, seed=1, max_plies=2)
    assert len(games2) == 8 and all(len(g["actions"]) == 2 for g in games2)


def test_selfplay_records_equal_packed_dict_histories():
    """sim.selfplay_records writes the packed training format directly (vectorised, no per-move Python objects): it must be
    records.pack_records of what selfplay_batch returns for the same seed -- states, CSR policy targets, outcomes, row order --
    including games that end inside the run (short games from a mating net)"""
    from sigma_zero_b200 import records
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.sim import selfplay_batch, selfplay_records
    torch.manual_seed(0)
    model = policyNN({}).eval()
    args = {"C": 2, "num_searches": 12}
    for c960, n, plies in ((True, 7, 9), (False, 3, 5)):
        games, c0 = selfplay_batch(model, args, n, c960=c960, seed=11, max_plies=plies)
        want = records.pack_records(games)
        got, c1 = selfplay_records(model, args, n, c960=c960, seed=11, max_plies=plies)
        assert c0["plies"] == c1["plies"]
        for k in ("states", "pi_index", "pi_prob", "pi_off", "z", "colour", "game"):
            assert np.array_equal(want[k], got[k]), k
        assert got["result"].shape == (n,) and set(got["result"].tolist()) <= {2, 1, 0, -1}
        dense = records.dense_policy(got, [0, len(got["z"]) - 1])
        assert np.allclose(dense.sum(1), 1.0, atol=1e-6)


def test_supervised_pgn_to_records_matches_oracle_replay():
    """PGN games replayed in lock step on the GPU engine -> packed one-hot records; every state, move index,
    colour and reward must equal what the reference's loop produces (generate_training_supervised.py:60-95, restated over the
    oracle), and the byte view of the packed planes must be the reference's compressed tensor (:91)"""
    from sigma_zero_b200.supervised import compressed_states, pgn_to_records, read_pgn, resolve_san, select_balanced
    games = select_balanced(read_pgn(util.PGN_SAMPLE))
    rec = pgn_to_records(games, strict=True)
    want_states, want_idx, want_z, want_col, want_game = [], [], [], [], []
    for gi, (h, san) in enumerate(games):
        _, ucis = util.replay_san_on_oracle(san, resolve_san)
        og = util.oracle_game(False, 518)
        r = {"1-0": 1, "0-1": -1, "1/2-1/2": 0}[h["Result"]]
        import chess
        for ply, u in enumerate(ucis):
            want_states.append(og.get_representation())
            want_idx.append(ref_path.move_to_index(chess.Move.from_uci(u), og.board.turn))
            want_col.append(bool(og.board.turn))
            want_z.append(r if ply % 2 == 0 else -r)
            want_game.append(gi)
            og.move_piece(chess.Move.from_uci(u))
    n = len(want_z)
    assert n == 33 + 16 + 4 and len(rec["z"]) == n
    assert np.array_equal(rec["states"], np.stack([hash_eval.pack_planes(s) for s in want_states]))
    assert rec["pi_index"].tolist() == want_idx and rec["z"].tolist() == want_z
    assert rec["colour"].tolist() == want_col and rec["game"].tolist() == want_game
    assert np.array_equal(rec["pi_off"], np.arange(n + 1)) and (rec["pi_prob"] == 1).all()
    ref = torch.stack([(torch.from_numpy(s.astype(np.uint8)) << torch.arange(8).view(1, 1, 8)).sum(dim=-1).to(torch.uint8) for s in want_states])
    assert np.array_equal(compressed_states(rec), ref.numpy())


def test_train_rl_main_loop_and_resume(tmp_path):
    """train_RL.main: the reference's outer loop (self-play -> save games -> fine-tune -> save weights and optimiser,
    train_RL.py:205-264) on the engine, and a resume from the files it wrote"""
    from sigma_zero_b200 import records
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.train_RL import main
    torch.manual_seed(0)
    args = {"C": 2, "num_searches": 8, "num_epochs": 3, "batch_size": 4, "chess960": True}
    logs = []
    model, hist = main(args, model=policyNN({}), num_games=4, out_dir=str(tmp_path), max_plies=3, passes_per_epoch=1,
                       train_device="cuda", log=logs.append)
    assert [h["epoch"] for h in hist] == [1, 2] and all(h["positions"] == 12 for h in hist)
    assert all(np.isfinite(np.array(h["losses"])).all() and len(h["losses"]) == 3 for h in hist)
    for epoch in (1, 2):
        rec = records.load(str(tmp_path / "games" / ("RL_960_%d.npz" % epoch)))
        assert rec["states"].shape == (12, 119) and sorted(set(rec["game"].tolist())) == [0, 1, 2, 3]
        assert (tmp_path / "saves" / ("RL_960_%d.pt" % epoch)).exists() and (tmp_path / "saves" / ("RL_960_opt_%d.pt" % epoch)).exists()
    saved = torch.load(str(tmp_path / "saves" / "RL_960_2.pt"), map_location="cpu")
    assert all(torch.equal(v.cpu(), saved[k]) for k, v in model.state_dict().items())
    # the CUDA trainer's state travels in torch's formats: 3 batches per epoch -> Adam step 6, StepLR counter 6
    opt2 = torch.load(str(tmp_path / "saves" / "RL_960_opt_2.pt"), map_location="cpu")
    assert len(opt2["state"]) == 129 and float(opt2["state"][0]["step"]) == 6 and opt2["state"][0]["exp_avg"].shape == saved["conv1.weight"].shape
    assert torch.load(str(tmp_path / "saves" / "RL_960_sched_2.pt"), map_location="cpu")["last_epoch"] == 6
    # resume: epoch 3 starts from the weights of epoch 2
    logs.clear()
    model2, hist2 = main(dict(args, start_epoch=3, num_epochs=4), model=policyNN({}), num_games=4, out_dir=str(tmp_path),
                         max_plies=2, passes_per_epoch=1, train_device="cuda", log=logs.append)
    assert [h["epoch"] for h in hist2] == [3] and not any("No saved weights" in str(l) for l in logs)
    assert (tmp_path / "saves" / "RL_960_3.pt").exists()
    opt3 = torch.load(str(tmp_path / "saves" / "RL_960_opt_3.pt"), map_location="cpu")
    assert float(opt3["state"][0]["step"]) == 6 + len(hist2[0]["losses"])          # the step counter continued from the checkpoint
