"""Host-side logic of the drop-in facade (no GPU): the closed-form move <-> index codec against the golden
vectors recorded from the UNMODIFIED reference codec, plane packing, game sharding, the flat weight buffer and
its broadcast over a world_size-2 gloo group, and bench.py's reference-arm JSON contract."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests import util


def test_facade_codec_known_answers():
    from sigma_zero_b200 import chess_compat as cc
    from sigma_zero_b200.chess_tensor import index_move, move_index
    kat = [("e2e4", 1, 116), ("e7e5", 0, 115), ("g1f3", 1, 4094), ("g8f6", 0, 3641), ("e1g1", 1, 1020),
           ("e1h1", 1, 1084), ("a7a8q", 1, 8), ("a7a8n", 1, 4104), ("b7a8r", 1, 4617), ("h2h1b", 0, 4296),
           ("g2h1n", 0, 4233)]
    for u, w, idx in kat:
        m = cc.Move.from_uci(u)
        assert move_index(m, bool(w)) == idx, u
        back = index_move(idx, bool(w), {u: True} if u.endswith("q") else None)
        assert back.uci() == u, (u, back.uci())


def test_facade_codec_golden(golden_dir):
    """every legal move of 333 positions reached by the reference: facade index == reference index, and decode inverts"""
    from sigma_zero_b200 import chess_compat as cc
    from sigma_zero_b200.chess_tensor import actionsToTensor, index_move, move_index, tensorToAction
    z = np.load(os.path.join(golden_dir, "codec.npz"))
    off = np.concatenate([[0], np.cumsum(z["idx_len"])])
    for i in range(0, len(z["sid"]), 3):
        ucis = str(z["uci"][i]).split()
        want = z["idx_flat"][off[i]:off[i + 1]].astype(int).tolist()
        white = (len(str(z["moves"][i]).split()) % 2) == 0
        moves = [cc.Move.from_uci(u) for u in ucis]
        got = sorted(move_index(m, white) for m in moves)
        assert got == sorted(want), i
        mask, qp = actionsToTensor(moves, white)
        assert mask.sum().item() == len(moves) and set(mask.nonzero().flatten().tolist()) == set(want)
        decoded = tensorToAction(mask, white, qp)
        assert sorted(m.uci() for m in decoded) == sorted(ucis), i
        for m in moves:
            assert index_move(move_index(m, white), white, qp).uci() == m.uci()


def test_plane_packing_roundtrip(golden_dir):
    from sigma_zero_b200 import runtime
    z = np.load(os.path.join(golden_dir, "codec.npz"))
    words = z["planes"][:16]
    planes = runtime.unpack_planes(words)
    assert planes.shape == (16, 119, 8, 8) and planes.dtype == bool
    assert np.array_equal(runtime.pack_planes(planes), words)
    # bit (row*8+col) of word k is planes[k,row,col]
    k, r, c = 5, 7, 4
    assert bool((int(words[0, k]) >> (r * 8 + c)) & 1) == bool(planes[0, k, r, c])


def test_shards_cover_all_games():
    from sigma_zero_b200.train_RL import shard_of
    for n in (1, 7, 500, 1024):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_of(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_state_dict_layout_and_flat_buffer():
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.train_RL import flatten_state_dict, unflatten_into
    torch.manual_seed(3)
    a = policyNN({})
    sd = a.state_dict()
    assert len(sd) == 252 and sum(v.numel() for k, v in sd.items() if not k.endswith("num_batches_tracked")) == 22809420 + 2 * (256 * 40 + 1)
    keys, flat = flatten_state_dict(sd)
    b = policyNN({})
    unflatten_into(b, keys, flat)
    for k in keys:
        assert torch.equal(a.state_dict()[k], b.state_dict()[k]), k
    if not torch.cuda.is_available():
        from sigma_zero_b200.engine import SzbError
        with pytest.raises(SzbError):                        # the self-play (eval) forward never falls back to torch
            policyNN({}).eval()(torch.zeros(1, 119, 8, 8))


def _bcast_worker(rank, world, port, out):
    import torch.distributed as dist
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.train_RL import broadcast_weights, flatten_state_dict, shard_of
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                    # ranks start with DIFFERENT weights
    m = policyNN({})
    nbytes = broadcast_weights(m, src=0)
    _, flat = flatten_state_dict(m.state_dict())
    digest = torch.tensor([float(flat.double().sum()), float(flat.abs().double().sum())], dtype=torch.float64)
    gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, digest)
    lo, hi = shard_of(500, rank, world)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo]))
    if rank == 0:
        out.put((nbytes, [g.tolist() for g in gathered], [int(s) for s in sizes]))
    dist.destroy_process_group()


def test_weight_broadcast_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    nbytes, digests, sizes = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert nbytes == (22809420 + 2 * (256 * 40 + 1)) * 4        # parameters + BN running statistics, fp32
    assert digests[0] == digests[1]                              # rank 1 now holds rank 0's weights
    assert sum(sizes) == 500


def test_bench_reference_arm_contract():
    out = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-sims", "8"], capture_output=True, text=True, timeout=600, cwd=util.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mcts_simulations_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    for k in ("unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config"):
        assert k in line


def test_record_format_roundtrip(tmp_path, golden_dir):
    """packed self-play records: states, sparse visit distributions and outcomes survive pack -> save -> load -> unpack"""
    from sigma_zero_b200 import chess_compat as cc, records, runtime
    from sigma_zero_b200.chess_tensor import actionsToTensor
    z = np.load(os.path.join(golden_dir, "codec.npz"))
    games = []
    for g in range(3):
        h = {"states": [], "actions": [], "rewards": [], "colours": []}
        for k in range(4):
            i = 7 * g + k
            white = (len(str(z["moves"][i]).split()) % 2) == 0
            ucis = str(z["uci"][i]).split()
            if not ucis:
                continue
            p = np.random.default_rng(i).random(len(ucis))
            p /= p.sum()
            h["states"].append(torch.from_numpy(runtime.unpack_planes(z["planes"][i])))
            h["actions"].append({cc.Move.from_uci(u): float(x) for u, x in zip(ucis, p)})
            h["colours"].append(white)
            h["rewards"].append((-1) ** k)
        games.append(h)
    rec = records.pack_records(games)
    n = sum(len(h["actions"]) for h in games)
    assert rec["states"].shape == (n, 119) and rec["pi_off"][-1] == len(rec["pi_index"]) and len(rec["z"]) == n
    records.save(str(tmp_path / "r.npz"), rec)
    back = records.load(str(tmp_path / "r.npz"))
    planes = records.unpack_states(back)
    dense = records.dense_policy(back)
    i = 0
    for g, h in enumerate(games):
        for s, a, r, c in zip(h["states"], h["actions"], h["rewards"], h["colours"]):
            assert np.array_equal(planes[i], s.numpy()) and back["z"][i] == r and back["colour"][i] == c and back["game"][i] == g
            want, _ = actionsToTensor(a, c)                      # the reference's own dense soft target (train_RL.py:43-45)
            assert np.allclose(dense[i], want.numpy(), atol=1e-7)
            i += 1


def test_training_path_matches_reference_architecture_and_learns():
    """the torch training forward of the facade's policyNN (model.train()) is the reference architecture:
    same outputs as the restated reference network for the same state_dict; a few fine-tuning steps on packed records reduce
    the reference's loss (train_RL.py:103-113)"""
    from oracle import ref_path
    from sigma_zero_b200 import records, runtime
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.train_RL import make_optimiser, train_on_records
    torch.manual_seed(0)
    model = policyNN({})
    ref = ref_path.build_policy_nn()
    ref.load_state_dict(model.state_dict())
    z = np.load(os.path.join(util.GOLDEN if hasattr(util, "GOLDEN") else os.path.join(util.ROOT, "tests", "golden"), "codec.npz"))
    planes = runtime.unpack_planes(z["planes"][:6])
    x = torch.from_numpy(planes).float()
    model.eval(); ref.eval()
    with torch.no_grad():
        p0, v0 = model.forward_torch(x, inference=True)
        p1, v1 = ref(x, inference=True)
    assert torch.equal(p0, p1) and torch.equal(v0, v1)
    # records: 6 positions, a peaked target on one legal move each, alternating outcomes
    rec = {"states": z["planes"][:6].astype(np.uint64), "pi_index": np.array([116, 115, 4094, 3641, 116, 115], dtype=np.uint16),
           "pi_prob": np.ones(6, dtype=np.float32), "pi_off": np.arange(7, dtype=np.int64),
           "z": np.array([1, -1, 1, -1, 0, 0], dtype=np.int8), "colour": np.array([1, 0, 1, 0, 1, 0], dtype=bool),
           "game": np.zeros(6, dtype=np.int32)}
    before = [t.clone() for t in model.parameters()]
    opt, sched = make_optimiser(model, lr=1e-3)
    hist = train_on_records(model, rec, epochs=6, batch_size=6, optimiser=opt, lr_scheduler=sched, device="cpu")
    assert len(hist) == 6 and sum(hist[-1]) < sum(hist[0])
    assert not model.training and any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))


def test_concat_records_rebases_csr_offsets():
    """record sets of consecutive game blocks (one per rank) concatenate into one set: CSR offsets re-based, rows kept"""
    from sigma_zero_b200 import records
    from sigma_zero_b200.train_RL import concat_records
    rng = np.random.default_rng(0)

    def fake(n, game0):
        lens = rng.integers(1, 6, n)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        idx = np.concatenate([np.zeros(0, np.int64)] + [np.sort(rng.choice(4672, k, replace=False)) for k in lens]).astype(np.uint16)
        prob = np.concatenate([np.zeros(0, np.float32)] + [np.full(k, 1.0 / k, np.float32) for k in lens])
        return {"states": rng.integers(0, 2 ** 63, (n, 119), dtype=np.uint64), "pi_index": idx, "pi_prob": prob, "pi_off": off,
                "z": rng.integers(-1, 2, n).astype(np.int8), "colour": rng.integers(0, 2, n).astype(bool),
                "game": (game0 + np.arange(n) // 2).astype(np.int32), "result": np.zeros((n + 1) // 2, np.int8)}
    a, b, c = fake(5, 0), fake(0, 3), fake(4, 3)
    whole = concat_records([a, b, c])
    assert len(whole["z"]) == 9 and whole["pi_off"][0] == 0 and whole["pi_off"][-1] == len(whole["pi_index"])
    assert np.all(np.diff(whole["pi_off"]) > 0)
    d = records.dense_policy(whole)
    assert np.array_equal(d[:5], records.dense_policy(a)) and np.array_equal(d[5:], records.dense_policy(c))
    assert np.array_equal(whole["states"][5:], c["states"]) and len(whole["result"]) == 3 + 0 + 2


def test_supervised_pgn_reader_and_san_resolver():
    """SURVEY 8f rank 3, host side: PGN text -> (headers, SAN main line) with comments / variations / NAGs dropped, the
    reference's balanced game selection, and SAN resolved against a legal-move list (here the oracle's) -- castling both ways,
    disambiguation, en passant, promotion by capture, mate"""
    from sigma_zero_b200.supervised import read_pgn, resolve_san, select_balanced
    games = list(read_pgn(util.PGN_SAMPLE))
    assert [h["Result"] for h, _ in games] == ["1-0", "1/2-1/2", "0-1", "*"]
    assert len(games[0][1]) == 33 and games[0][1][21] == "Nbd7" and games[0][1][22] == "O-O-O" and games[0][1][-1] == "Rd8#"
    assert games[1][1] == ["e4", "a6", "e5", "d5", "exd6", "Nf6", "dxc7", "Nc6", "cxd8=N", "g6", "Nf3", "Bg7", "Bc4", "O-O", "O-O", "Rxd8"]
    assert list(read_pgn(util.PGN_SAMPLE.encode())) == games
    g, ucis = util.replay_san_on_oracle(games[0][1], resolve_san)
    assert ucis[:3] == ["e2e4", "e7e5", "g1f3"] and ucis[22] == "e1c1" and ucis[21] == "b8d7"
    assert g.board.outcome() is not None and g.board.result() == "1-0"
    g, ucis = util.replay_san_on_oracle(games[1][1], resolve_san)
    assert ucis[4] == "e5d6" and ucis[8] == "c7d8n" and ucis[13] == "e8g8" and ucis[14] == "e1g1"
    g, _ = util.replay_san_on_oracle(games[2][1], resolve_san)
    assert g.board.result() == "0-1"
    with pytest.raises(ValueError):
        util.replay_san_on_oracle(["e4", "e5", "Ke3"], resolve_san)
    # selection: one white win + one black win + the draw are balanced; asking for 2 stops the scan after the first two games
    assert [h["Result"] for h, _ in select_balanced(games)] == ["1-0", "1/2-1/2", "0-1"]
    assert [h["Result"] for h, _ in select_balanced(games, num_games=1)] == ["1/2-1/2"]


def _gather_worker(rank, world, port, out):
    import torch.distributed as dist
    from sigma_zero_b200 import records
    from sigma_zero_b200.train_RL import concat_records, shard_of
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_of(7, rank, world)                       # 7 games over 2 ranks: blocks of 4 and 3
    rng = np.random.default_rng(40 + rank)
    n = 3 * (hi - lo)                                       # three positions per game
    lens = rng.integers(1, 5, n)
    rec = {"states": rng.integers(0, 2 ** 63, (n, 119), dtype=np.uint64),
           "pi_index": np.concatenate([np.sort(rng.choice(4672, k, replace=False)) for k in lens]).astype(np.uint16),
           "pi_prob": np.concatenate([np.full(k, 1.0 / k, np.float32) for k in lens]),
           "pi_off": np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), "z": rng.integers(-1, 2, n).astype(np.int8),
           "colour": (np.arange(n) % 2 == 0), "game": (lo + np.arange(n) // 3).astype(np.int32),
           "result": np.zeros(hi - lo, np.int8)}
    parts = [None] * world
    dist.all_gather_object(parts, rec)                      # what train_RL.rl_iteration does with every rank's shard
    whole = concat_records(parts)
    digest = (len(whole["z"]), int(whole["pi_off"][-1]), whole["game"].tolist(), float(records.dense_policy(whole).sum()),
              int(whole["states"].sum(dtype=np.uint64) % np.uint64(1 << 62)))
    out.put((rank, digest))
    dist.destroy_process_group()


def test_record_gather_gloo_world2():
    """the N > 1 host path of an iteration: every rank's packed records gathered and concatenated -- both ranks end
    up with the same record set, games in global order"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]
    n, m, games, total, _ = got[0]
    assert n == 21 and games == sorted(games) and games[0] == 0 and games[-1] == 6 and abs(total - 21.0) < 1e-4


def test_product_arm_has_no_cpu_fallback():
    """without a CUDA device the bench's own arm must stop loudly (the product path never routes through the oracle)"""
    if torch.cuda.is_available():
        return
    for extra in ([], ["--workload", "c4"]):
        out = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--steps", "1", "--warmup", "1"] + extra,
                             capture_output=True, text=True, timeout=600, cwd=util.ROOT)
        assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
