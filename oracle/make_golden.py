"""Generates tests/golden/*.npz by running the UNMODIFIED reference files from /root/reference.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (the GPU box has no /root/reference); the
fixtures it writes are committed.  The reference modules mcts.py, mctsnode.py, chess_tensor.py, network.py
and sim.py are imported as-is, with oracle/chess standing in for python-chess (not installable here).
For every case the restatement in oracle/ref_path.py is run on the same inputs and must agree exactly
before the fixture is written -- that is what pins the oracle.

    python -m oracle.make_golden            # from the repo root
"""
import contextlib
import hashlib
import io
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))     # `chess` stand-in
REF = "/root/reference"
sys.path.insert(1, REF)

import chess  # noqa: E402
from oracle import hash_eval, ref_path  # noqa: E402

import chess_tensor as ref_ct  # noqa: E402  (reference)
import mcts as ref_mcts  # noqa: E402        (reference)
import network as ref_net  # noqa: E402      (reference)
import sim as ref_sim  # noqa: E402          (reference)

OUT = os.path.join(ROOT, "tests", "golden")


class FakeModel(torch.nn.Module):
    """hash evaluator dressed as the reference's model: model(x[1,119,8,8], inference=True) -> (policy, value)"""

    def forward(self, x, inference=False):
        p, v = hash_eval.evaluator(x[0].cpu().numpy() != 0)
        return torch.from_numpy(p).unsqueeze(0), torch.tensor([[float(v)]], dtype=torch.float32)


def ref_game(c960: bool, start_id: int, moves):
    """reference ChessTensor advanced through `moves` (uci)."""
    if c960:
        st = random.getstate()
        # find a seed whose first randint(0, 959) is start_id -- cheaper: patch the call
        orig = random.randint
        random.randint = lambda a, b: start_id
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                g = ref_ct.ChessTensor(chess960=True)
        finally:
            random.randint = orig
            random.setstate(st)
    else:
        g = ref_ct.ChessTensor()
    for u in moves:
        g.move_piece(chess.Move.from_uci(u))
    return g


def port_game(c960: bool, start_id: int, moves):
    g = ref_path.RefGame(chess960=c960, start_id=start_id)
    for u in moves:
        g.move_piece(chess.Move.from_uci(u))
    return g


def random_playout(rng, c960, start_id, max_plies=400):
    b = chess.Board.from_chess960_pos(start_id) if c960 else chess.Board()
    moves = []
    while b.outcome() is None and len(moves) < max_plies:
        legal = list(b.legal_moves)
        m = legal[rng.integers(len(legal))]
        moves.append(m.uci())
        b.push(m)
    return moves, b.result()


def gen_codec(rng):
    """positions along random playouts: planes + sorted legal indices from the reference codec."""
    recs = []
    for gi in range(24):
        c960 = gi % 2 == 1
        sid = int(rng.integers(960)) if c960 else 518
        moves, result = random_playout(rng, c960, sid)
        g = ref_game(c960, sid, [])
        p = port_game(c960, sid, [])
        sample = set(rng.choice(len(moves) + 1, size=min(12, len(moves) + 1), replace=False).tolist()) | {0, len(moves)}
        for ply in range(len(moves) + 1):
            if ply in sample:
                white = g.board.turn
                legal = list(g.board.legal_moves)
                mask, qp = ref_ct.actionsToTensor(legal, white)
                idx = mask.nonzero().flatten().numpy().astype(np.int32)
                assert mask.max() <= 1 and len(idx) == len(legal)
                decoded = ref_ct.tensorToAction(mask, white, qp) if len(legal) else []
                assert set(m.uci() for m in decoded) == set(m.uci() for m in legal)
                planes = g.get_representation().numpy()
                # the restatement must agree
                pm, pq = ref_path.legal_mask(legal, white)
                assert np.array_equal(pm, mask.numpy())
                assert [ref_path.index_to_move(i, white, pq).uci() for i in idx] == [m.uci() for m in decoded]
                assert np.array_equal(p.get_representation(), planes), (gi, ply)
                v, t = g.get_value_and_terminated()
                assert (v, t) == p.value_and_terminated()
                recs.append(dict(c960=c960, sid=sid, moves=" ".join(moves[:ply]), idx=idx,
                                 planes=hash_eval.pack_planes(planes), value=v, terminal=t,
                                 uci=" ".join(m.uci() for m in decoded)))
            if ply < len(moves):
                g.move_piece(chess.Move.from_uci(moves[ply]))
                p.move_piece(chess.Move.from_uci(moves[ply]))
    np.savez_compressed(
        os.path.join(OUT, "codec.npz"),
        c960=np.array([r["c960"] for r in recs]), sid=np.array([r["sid"] for r in recs], dtype=np.int32),
        moves=np.array([r["moves"] for r in recs]), uci=np.array([r["uci"] for r in recs]),
        idx_flat=np.concatenate([r["idx"] for r in recs]).astype(np.int16),
        idx_len=np.array([len(r["idx"]) for r in recs], dtype=np.int32),
        planes=np.stack([r["planes"] for r in recs]),
        value=np.array([r["value"] for r in recs], dtype=np.int8),
        terminal=np.array([r["terminal"] for r in recs]))
    print("codec: %d positions" % len(recs))


def gen_search(rng):
    """visit counts of the unmodified MCTS0.search driven by the hash evaluator."""
    model = FakeModel()
    cases = []
    specs = []
    for gi in range(10):
        c960 = gi % 2 == 1
        sid = int(rng.integers(960)) if c960 else 518
        moves, _ = random_playout(rng, c960, sid)
        for back in (len(moves), 30, 6, 3, 1):     # start, middle, and positions a few plies before the end
            ply = max(0, len(moves) - back)
            specs.append((c960, sid, moves[:ply]))
    for ci, (c960, sid, moves) in enumerate(specs):
        learning = ci % 2 == 0
        n_search = (40, 120, 200)[ci % 3]
        C = (2, 1.5)[ci % 4 == 3]
        g = ref_game(c960, sid, moves)
        if g.board.is_game_over():
            continue
        args = {"C": C, "num_searches": n_search}
        probs = ref_mcts.MCTS0(game=g, args=args, model=model).search(g.board, verbose=False, learning=learning)
        white = g.board.turn
        total = n_search - 1
        idx = np.array([ref_path.move_to_index(m, white) for m in probs], dtype=np.int32)
        visits = np.array([round(v * total) for v in probs.values()], dtype=np.int32)
        assert list(idx) == sorted(idx) and visits.sum() == total
        # the restatement must agree exactly
        pp, _ = ref_path.search(port_game(c960, sid, moves), n_search, C, hash_eval.evaluator, learning=learning)
        assert [m.uci() for m in pp] == [m.uci() for m in probs], ci
        assert list(pp.values()) == list(probs.values()), ci
        cases.append(dict(c960=c960, sid=sid, moves=" ".join(moves), learning=learning, n=n_search, C=C,
                          idx=idx, visits=visits, uci=" ".join(m.uci() for m in probs)))
        print("search case %d: ply %d n=%d learning=%s children=%d max=%d" % (
            ci, len(moves), n_search, learning, len(idx), visits.max()))
    np.savez_compressed(
        os.path.join(OUT, "search.npz"),
        c960=np.array([c["c960"] for c in cases]), sid=np.array([c["sid"] for c in cases], dtype=np.int32),
        moves=np.array([c["moves"] for c in cases]), learning=np.array([c["learning"] for c in cases]),
        n=np.array([c["n"] for c in cases], dtype=np.int32), C=np.array([c["C"] for c in cases], dtype=np.float64),
        idx_flat=np.concatenate([c["idx"] for c in cases]).astype(np.int16),
        visits_flat=np.concatenate([c["visits"] for c in cases]).astype(np.int32),
        idx_len=np.array([len(c["idx"]) for c in cases], dtype=np.int32),
        uci=np.array([c["uci"] for c in cases]))


def gen_selfplay():
    """sim.play_game (unmodified) with the hash evaluator; the restatement must replay it move for move."""
    model = FakeModel()
    games = []
    for gi, (c960, n_search) in enumerate([(False, 12), (True, 8)]):
        args = {"C": 2, "num_searches": n_search}
        random.seed(100 + gi)
        sid = random.randint(0, 959) if c960 else 518
        random.seed(100 + gi)
        np.random.seed(7 + gi)
        with contextlib.redirect_stdout(io.StringIO()):
            h = ref_sim.play_game(model, args, c960=c960)
        # replay with the restatement: identical global numpy RNG stream
        np.random.seed(7 + gi)
        hp = ref_path.play_game(hash_eval.evaluator, args, c960=c960, start_id=sid)
        ref_moves = []
        g = ref_game(c960, sid, [])
        # recover the moves the reference played from consecutive states is awkward; use action dict + states length
        assert len(h["actions"]) == len(hp["actions"]) and h["rewards"] == hp["rewards"], gi
        for a, b in zip(h["actions"], hp["actions"]):
            assert [m.uci() for m in a] == [m.uci() for m in b] and list(a.values()) == list(b.values())
        for s, t in zip(h["states"], hp["states"]):
            assert np.array_equal(s.numpy(), t)
        games.append(dict(c960=c960, sid=sid, n=n_search, plies=len(h["actions"]), rewards=h["rewards"],
                          result=hp["result"],
                          first_state=hash_eval.pack_planes(h["states"][0].numpy()),
                          last_state=hash_eval.pack_planes(h["states"][-1].numpy())))
        print("selfplay game %d: %d plies result %s" % (gi, len(h["actions"]), hp["result"]))
        del ref_moves, g
    np.savez_compressed(
        os.path.join(OUT, "selfplay.npz"),
        c960=np.array([g["c960"] for g in games]), sid=np.array([g["sid"] for g in games], dtype=np.int32),
        n=np.array([g["n"] for g in games], dtype=np.int32), plies=np.array([g["plies"] for g in games], dtype=np.int32),
        result=np.array([g["result"] for g in games]),
        last_state=np.stack([g["last_state"] for g in games]), first_state=np.stack([g["first_state"] for g in games]))


def gen_network(rng):
    """seeded init equality (reference policyNN vs restatement) + fp32 CPU outputs on three positions."""
    torch.manual_seed(0)
    ref = ref_net.policyNN({}).eval()
    torch.manual_seed(0)
    port = ref_path.build_policy_nn().eval()
    sd_r, sd_p = ref.state_dict(), port.state_dict()
    assert list(sd_r.keys()) == list(sd_p.keys()) and len(sd_r) == 252
    h = hashlib.sha256()
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_p[k]), k
        h.update(sd_r[k].numpy().tobytes())
    xs = []
    for moves in ([], ["e2e4"], ["e2e4", "e7e5", "g1f3", "b8c6", "f1b5", "a7a6"]):
        xs.append(ref_game(False, 518, moves).get_representation().float())
    x = torch.stack(xs)
    with torch.no_grad():
        pr, vr = ref(x, inference=True)
        pp, vp = port(x, inference=True)
    assert torch.equal(pr, pp) and torch.equal(vr, vp)
    np.savez_compressed(os.path.join(OUT, "network.npz"), sha256=np.array(h.hexdigest()),
                        planes=np.stack([hash_eval.pack_planes(t.numpy()) for t in xs]),
                        policy=pr.numpy(), value=vr.numpy(), n_keys=np.array(len(sd_r)),
                        n_params=np.array(sum(p.numel() for p in ref.parameters())))
    print("network: sha256", h.hexdigest()[:16], "value", vr.flatten().tolist())


def gen_numerics(rng):
    """torch-generated vectors for PUCT / cascade-sum / noise (Appendix C/D of SURVEY.md)."""
    import math
    xs, sums = [], []
    for t in range(64):
        logits = (rng.normal(size=4672) * 3).astype(np.float32)
        sm = torch.softmax(torch.from_numpy(logits), 0).numpy()
        mask = np.zeros(4672, np.float32)
        mask[rng.choice(4672, int(rng.integers(1, 80)), replace=False)] = 1
        x = (sm * mask).astype(np.float32) if t % 4 else rng.random(4672).astype(np.float32)
        xs.append(x)
        sums.append(torch.sum(torch.from_numpy(x)).item())
    pn, pw, pp, pN, pc, pu, pl = [], [], [], [], [], [], []
    for t in range(400):
        c = int(rng.integers(1, 80))
        n = rng.integers(0, 60, size=c)
        w = np.where(n == 0, 0.0, rng.normal(size=c) * n * 0.3)
        p = rng.random(c).astype(np.float32)
        Np = int(n.sum() + 1)
        C = (2, 1.5, 4)[t % 3]
        vc = torch.tensor([int(v) for v in n])
        vsum = torch.tensor([float(v) for v in w])
        prior = torch.tensor([float(v) for v in p])
        ucb = (1 - (vsum / (vc + 1e-6) + 1) / 2) + C * (math.sqrt(Np) / (vc + 1)) * prior
        pad = lambda a, dt: np.pad(np.asarray(a, dtype=dt), (0, 80 - c))
        pn.append(pad(n, np.int32)); pw.append(pad(w, np.float64)); pp.append(pad(p, np.float32))
        pN.append(Np); pc.append(C); pu.append(pad(ucb.numpy(), np.float32)); pl.append(c)
    np.savez_compressed(os.path.join(OUT, "numerics.npz"), sum_x=np.stack(xs), sum_y=np.array(sums, dtype=np.float32),
                        puct_n=np.stack(pn), puct_w=np.stack(pw), puct_p=np.stack(pp), puct_N=np.array(pN, dtype=np.int32),
                        puct_C=np.array(pc, dtype=np.float64), puct_ucb=np.stack(pu),
                        puct_len=np.array(pl, dtype=np.int32))
    print("numerics written")


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)
    gen_numerics(rng)
    gen_codec(rng)
    gen_network(rng)
    gen_search(rng)
    gen_selfplay()


if __name__ == "__main__":
    main()
