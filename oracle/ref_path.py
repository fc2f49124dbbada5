"""CPU restatement of the reference self-play hot path (TEST INFRASTRUCTURE ONLY -- oracle).

What is restated, and from where (all cites into /root/reference):
  RefGame            <- chess_tensor.py:30-178  ChessTensor (board + 8-ply plane history, both perspectives)
  move_to_index      <- chess_tensor.py:221-306 actionToTensor
  index_to_move      <- chess_tensor.py:309-410 tensorToAction
  legal_mask         <- chess_tensor.py:190-218 actionsToTensor
  RefNode / search   <- mctsnode.py:7-63, mcts.py:39-122
  RefPolicyNN        <- network.py:36-192 (same state_dict keys, same construction order => same seeded init)
  play_game          <- sim.py:31-99

The arithmetic that decides visit counts is taken from oracle/numerics.py (one IEEE op at a time, pinned
against torch).  The chess rules come from the `chess` stand-in in oracle/chess (python-chess is not
installable here).  This module is validated against the UNMODIFIED reference files in this container by
oracle/make_golden.py (identical planes / indices / visit counts on seeded runs) and travels to the GPU
box, where /root/reference does not exist.

PARITY STATUS: pinned against the reference's own code run here (tests/golden/*.npz + make_golden.py);
the python-chess boundary below it is pinned only by external KATs ("parity unpinned" by the reference).
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)          # makes `import chess` resolve to oracle/chess
import chess  # noqa: E402  (the oracle stand-in)

from . import numerics  # noqa: E402

N_PLANES = 119
N_ACTIONS = 4672

# (dx, dy) in player-perspective board coordinates: x = column (+ right), y = row (+ down the tensor)
QUEEN_DIRS = [(0, -1), (1, -1), (1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1)]
KNIGHT_DIRS = [(1, -2), (2, -1), (2, 1), (1, 2), (-1, 2), (-2, 1), (-2, -1), (-1, -2)]


# --------------------------------------------------------------------------------------------------
# move <-> index codec
# --------------------------------------------------------------------------------------------------
def _rc(square: int, white: bool):
    """player-perspective (row, col) of an absolute square (chess_tensor.py:258-268)."""
    if white:
        return 7 - square // 8, square % 8
    return square // 8, 7 - square % 8


def _square(row: int, col: int, white: bool) -> int:
    if white:
        return (7 - row) * 8 + col
    return row * 8 + (7 - col)


def move_to_index(move, white: bool) -> int:
    row, col = _rc(move.from_square, white)
    to_row, to_col = _rc(move.to_square, white)
    dx, dy = to_col - col, to_row - row
    if dx == 0 or dy == 0 or abs(dx) == abs(dy):
        if move.promotion in (chess.KNIGHT, chess.BISHOP, chess.ROOK):
            plane = 64 + 3 * (move.promotion - chess.KNIGHT) + (0 if dx == 0 else (1 if dx > 0 else 2))
        else:
            direction = QUEEN_DIRS.index(((dx > 0) - (dx < 0), (dy > 0) - (dy < 0)))
            plane = direction * 7 + max(abs(dx), abs(dy)) - 1
    else:
        plane = 56 + KNIGHT_DIRS.index((dx, dy))
    return plane * 64 + row * 8 + col


def index_to_move(index: int, white: bool, queen_promotions=frozenset()):
    """queen_promotions: set of (from_square, to_square) of the legal queen promotions (the reference keeps a
    dict keyed by uci string, chess_tensor.py:205-209,370-371)."""
    plane, rest = divmod(int(index), 64)
    row, col = divmod(rest, 8)
    promotion = None
    if plane < 56:
        dx, dy = QUEEN_DIRS[plane // 7]
        dist = plane % 7 + 1
        to_row, to_col = row + dy * dist, col + dx * dist
    elif plane < 64:
        dx, dy = KNIGHT_DIRS[plane - 56]
        to_row, to_col = row + dy, col + dx
    else:
        k = plane - 64
        to_row = row - 1
        to_col = col + (0, 1, -1)[k % 3]
        promotion = (chess.KNIGHT, chess.BISHOP, chess.ROOK)[k // 3]
    f, t = _square(row, col, white), _square(to_row, to_col, white)
    if plane < 56 and (f, t) in queen_promotions:
        promotion = chess.QUEEN
    return chess.Move(f, t, promotion)


def legal_mask(moves, white: bool):
    """fp32[4672] 0/1 mask plus the queen-promotion set (actionsToTensor)."""
    mask = np.zeros(N_ACTIONS, dtype=np.float32)
    qp = set()
    for m in moves:
        if m.promotion == chess.QUEEN:
            qp.add((m.from_square, m.to_square))
        mask[move_to_index(m, white)] += np.float32(1.0)
    return mask, frozenset(qp)


# --------------------------------------------------------------------------------------------------
# game state with plane history
# --------------------------------------------------------------------------------------------------
def _piece_planes(board) -> np.ndarray:
    """bool[12,8,8]: white P N B R Q K then black, row = rank index, col = file (chess_tensor.py:38-63)."""
    out = np.zeros((12, 8, 8), dtype=bool)
    for sq in chess.SQUARES:
        pc = board.piece_at(sq)
        if pc:
            out[(pc.piece_type - 1) + (0 if pc.color else 6), sq // 8, sq % 8] = True
    return out


class RefGame:
    """Board + the last 8 positions' planes.  frames[0] is the current position."""

    T = 8

    def __init__(self, chess960: bool = False, start_id=None, board=None):
        if board is not None:
            self.board = board
        elif chess960:
            self.board = chess.Board.from_chess960_pos(random.randint(0, 959) if start_id is None else start_id)
        else:
            self.board = chess.Board()
        # (pieces12 absolute white-first, rep>=2, rep>=3)
        self.frames = [(_piece_planes(self.board), False, False)]
        # L planes: moved-flag, WK, WQ, BK, BQ, no-progress-flag.  Hard-coded at the start (chess_tensor.py:79-86)
        self.flags = (False, True, True, True, True, False)

    def copy(self) -> "RefGame":
        g = RefGame.__new__(RefGame)
        g.board = self.board.copy()
        g.frames = list(self.frames)
        g.flags = self.flags
        return g

    def move_piece(self, move) -> None:
        if move not in self.board.legal_moves:
            raise ValueError("Invalid move")
        b = self.board
        b.push(move)
        frame = (_piece_planes(b), b.is_repetition(2), b.is_repetition(3))
        self.frames = [frame] + self.frames[: self.T - 1]
        self.flags = (
            bool(len(b.move_stack)),
            b.has_kingside_castling_rights(chess.WHITE), b.has_queenside_castling_rights(chess.WHITE),
            b.has_kingside_castling_rights(chess.BLACK), b.has_queenside_castling_rights(chess.BLACK),
            bool(b.halfmove_clock),
        )

    def get_representation(self) -> np.ndarray:
        """bool[119,8,8] in the side-to-move's perspective (chess_tensor.py:131-142)."""
        white = self.board.turn
        out = np.zeros((N_PLANES, 8, 8), dtype=bool)
        for t, (pieces, r2, r3) in enumerate(self.frames):
            base = 14 * t
            if white:
                out[base: base + 12] = pieces
            else:
                out[base: base + 6] = pieces[6:12]
                out[base + 6: base + 12] = pieces[0:6]
            out[base + 12] = r2
            out[base + 13] = r3
        moved, wk, wq, bk, bq, noprog = self.flags
        out[112] = white
        out[113] = moved
        own = (wk, wq) if white else (bk, bq)
        opp = (bk, bq) if white else (wk, wq)
        out[114], out[115], out[116], out[117] = own[0], own[1], opp[0], opp[1]
        out[118] = noprog
        # White sees rank 8 on row 0 (row flip); Black sees a 180-degree turn of that = column flip of absolute
        return out[:, ::-1, :].copy() if white else out[:, :, ::-1].copy()

    def value_and_terminated(self):
        o = self.board.outcome()
        if o is None:
            return 0, False
        return (0 if o.winner is None else -1), True


# --------------------------------------------------------------------------------------------------
# search
# --------------------------------------------------------------------------------------------------
class RefNode:
    __slots__ = ("parent", "move", "index", "prior", "children", "n", "w", "game", "white")

    def __init__(self, parent=None, move=None, index=-1, prior=0.0, white=True):
        self.parent, self.move, self.index, self.prior, self.white = parent, move, index, prior, white
        self.children = []
        self.n = 0
        self.w = 0.0          # python double, like Node.value_sum
        self.game = None


def _select(node: RefNode, c) -> RefNode:
    ch = node.children
    best = numerics.puct_select([k.n for k in ch], [k.w for k in ch], [k.prior for k in ch], node.n, c)
    return ch[best]


def search(game: RefGame, num_searches: int, c, evaluator, learning: bool = False, trace=None):
    """Restates MCTS0.search.  evaluator(bool[119,8,8]) -> (softmax fp32[4672], value python-float/fp32).

    Returns (action_probs {Move: fraction}, root RefNode).  `trace`, if a list, receives one tuple per
    simulation: (leaf path as move indices, terminal flag, backed-up leaf value).
    """
    root = RefNode(white=bool(game.board.turn))
    root.game = game
    root.n = 1
    for _ in range(num_searches):
        node = root
        while node.children:
            node = _select(node, c)
        if node.parent is not None:
            node.game = node.parent.game.copy()
            node.game.move_piece(node.move)
        value, terminal = node.game.value_and_terminated()
        if not terminal:
            moves = list(node.game.board.legal_moves)
            mask, qp = legal_mask(moves, node.white)
            policy, v = evaluator(node.game.get_representation())
            p = numerics.normalise_policy(policy, mask)
            idx = np.nonzero(p)[0]
            priors = p[idx]
            if learning:
                priors = numerics.noisy_prior(priors)
            value = float(np.float32(v))
            for i, pr in zip(idx, priors):
                node.children.append(RefNode(node, index_to_move(i, node.white, qp), int(i), float(pr), not node.white))
        leaf_value = value
        if trace is not None:
            path, k = [], node
            while k.parent is not None:
                path.append(k.index)
                k = k.parent
            trace.append((tuple(reversed(path)), bool(terminal), float(leaf_value)))
        k = node
        while k is not None:
            k.w += value
            k.n += 1
            value = -value
            k = k.parent
    total = sum(ch.n for ch in root.children)
    probs = {ch.move: ch.n / total for ch in root.children}
    return probs, root


# --------------------------------------------------------------------------------------------------
# network (torch, CPU fp32)
# --------------------------------------------------------------------------------------------------
def build_policy_nn(in_channels: int = 119):
    """A torch module with network.py's architecture, state_dict keys and parameter construction order."""
    import torch
    import torch.nn as nn

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(256, 256, 3, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(256)
            self.relu = nn.ReLU(inplace=True)
            self.conv2 = nn.Conv2d(256, 256, 3, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(256)

        def forward(self, x):
            y = self.relu(self.bn1(self.conv1(x)))
            y = self.bn2(self.conv2(y))
            y += x
            return self.relu(y)

    class RefPolicyNN(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(in_channels, 256, 3, padding=1, bias=False)
            self.norm_layer = nn.BatchNorm2d(256)
            self.conv_p1 = nn.Conv2d(256, 256, 1, bias=False)
            self.p_norm1 = nn.BatchNorm2d(256)
            self.conv_p2 = nn.Conv2d(256, 73, 1)
            self.conv_v1 = nn.Conv2d(256, 1, 1, bias=False)
            self.v_norm = nn.BatchNorm2d(1)
            self.fc_v1 = nn.Linear(64, 256)
            self.fc_v2 = nn.Linear(256, 1)
            self.resnet_blocks = nn.Sequential(*[Block() for _ in range(19)])

        def forward(self, x, inference=False):
            x = torch.relu(self.norm_layer(self.conv1(x)))
            x = self.resnet_blocks(x)
            p = torch.relu(self.p_norm1(self.conv_p1(x)))
            p = torch.flatten(self.conv_p2(p), start_dim=1)
            v = torch.relu(self.v_norm(self.conv_v1(x)))
            v = torch.relu(self.fc_v1(torch.flatten(v, start_dim=1)))
            v = torch.tanh(self.fc_v2(v))
            if inference:
                p = torch.softmax(p, dim=1)
            return p, v

    return RefPolicyNN()


def torch_evaluator(model):
    """evaluator for search(): batch-1 fp32 forward on the model's device, like mcts.py:72-77."""
    import torch

    dev = next(model.parameters()).device

    def ev(planes):
        with torch.no_grad():
            x = torch.from_numpy(np.ascontiguousarray(planes)).float().unsqueeze(0).to(dev)
            p, v = model(x, inference=True)
        return p.squeeze(0).cpu().numpy(), v.item()

    return ev


# --------------------------------------------------------------------------------------------------
# self-play (sim.py:31-99)
# --------------------------------------------------------------------------------------------------
def play_game(evaluator, args, c960=False, start_id=None, max_plies=None, rng=None):
    """One self-play game.  Returns the reference's history dict plus counters.

    Move sampling uses `rng` (np.random.Generator) when given, else numpy's global RNG like sim.py:68.
    max_plies bounds the run for benchmarking (the reference has no cap)."""
    game = RefGame(chess960=c960, start_id=start_id)
    hist = {"states": [], "actions": [], "rewards": [], "colours": []}
    sims = 0
    while game.board.outcome() is None and (max_plies is None or len(hist["actions"]) < max_plies):
        state = game.get_representation()
        probs, _ = search(game, args["num_searches"], args["C"], evaluator, learning=True)
        sims += args["num_searches"]
        keys, vals = list(probs.keys()), list(probs.values())
        k = (rng.choice(len(keys), p=vals) if rng is not None else np.random.choice(len(keys), p=vals))
        hist["states"].append(state)
        hist["actions"].append(probs)
        hist["colours"].append(game.board.turn)
        game.move_piece(keys[int(k)])
    result = game.board.result()
    reward = 1 if result == "1-0" else (-1 if result == "0-1" else 0)
    hist["rewards"] = [reward if i % 2 == 0 else -reward for i in range(len(hist["actions"]))]
    hist["result"] = result
    hist["sims"] = sims
    return hist
