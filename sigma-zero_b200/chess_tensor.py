"""Drop-in for the reference's chess_tensor.py (ChessTensor, actionsToTensor, actionToTensor, tensorToAction).

Board state, legality, outcomes and the 119 input planes come from the CUDA engine (libszb200: szb_games_*,
szb_legal_moves, szb_encode); this file is the host-side object model plus the closed-form move <-> index
codec (reference chess_tensor.py:190-410).  One ChessTensor = one game; it is replayed into engine slot 0 when
it is touched (the batched self-play path in sim.py keeps thousands of games resident instead)."""
import random

import numpy as np
import torch

from . import chess_compat as chess
from . import runtime

N_ACTIONS = 4672
_QUEEN_DIRS = [(0, -1), (1, -1), (1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1)]
_KNIGHT_DIRS = [(1, -2), (2, -1), (2, 1), (1, 2), (-1, 2), (-2, 1), (-2, -1), (-1, -2)]


# ---- codec -----------------------------------------------------------------------------------------
def _view(square, white):
    """(row, col) of a square in the mover's view (reference :258-268)."""
    return (7 - square // 8, square % 8) if white else (square // 8, 7 - square % 8)


def move_index(move, color=True):
    row, col = _view(move.from_square, color)
    trow, tcol = _view(move.to_square, color)
    dx, dy = tcol - col, trow - row
    if dx == 0 or dy == 0 or abs(dx) == abs(dy):
        if move.promotion in (chess.KNIGHT, chess.BISHOP, chess.ROOK):
            plane = 64 + 3 * (move.promotion - chess.KNIGHT) + (0 if dx == 0 else (1 if dx > 0 else 2))
        else:
            plane = _QUEEN_DIRS.index(((dx > 0) - (dx < 0), (dy > 0) - (dy < 0))) * 7 + max(abs(dx), abs(dy)) - 1
    else:
        plane = 56 + _KNIGHT_DIRS.index((dx, dy))
    return plane * 64 + row * 8 + col


def index_move(index, color=True, queen_promotion=None):
    """Inverse of move_index.  queen_promotion: dict keyed by uci strings ending in 'q' (reference :205-209)."""
    plane, rest = divmod(int(index), 64)
    row, col = divmod(rest, 8)
    promo = None
    if plane < 56:
        dx, dy = _QUEEN_DIRS[plane // 7]
        d = plane % 7 + 1
        trow, tcol = row + dy * d, col + dx * d
    elif plane < 64:
        dx, dy = _KNIGHT_DIRS[plane - 56]
        trow, tcol = row + dy, col + dx
    else:
        k = plane - 64
        trow, tcol = row - 1, col + (0, 1, -1)[k % 3]
        promo = (chess.KNIGHT, chess.BISHOP, chess.ROOK)[k // 3]
    sq = (lambda r, c: (7 - r) * 8 + c) if color else (lambda r, c: r * 8 + (7 - c))
    f, t = sq(row, col), sq(trow, tcol)
    if plane < 56 and queen_promotion:
        if queen_promotion.get(chess.SQUARE_NAMES[f] + chess.SQUARE_NAMES[t] + "q", False):
            promo = chess.QUEEN
    return chess.make_move(f, t, promo)


def actionToTensor(move, color=chess.WHITE, prob=1):
    t = torch.zeros(N_ACTIONS)
    t[move_index(move, color)] = prob
    return t


def actionsToTensor(valid_moves, color=chess.WHITE):
    """mask (or probability vector when given a dict) over the 4672 actions + the queen-promotion dict"""
    t = torch.zeros(N_ACTIONS)
    queen_promotion = {}
    is_dict = isinstance(valid_moves, dict)
    for m in valid_moves:
        if m.promotion == chess.QUEEN:
            queen_promotion[str(m.uci())] = True
        t[move_index(m, color)] += valid_moves[m] if is_dict else 1
    return t, queen_promotion


def tensorToAction(moves, color=chess.WHITE, queen_promotion={}):
    return [index_move(int(i), color, queen_promotion) for i in torch.as_tensor(moves).nonzero().flatten()]


# ---- game object -----------------------------------------------------------------------------------
class ChessTensor:
    def __init__(self, chess960=False):
        self.M, self.T, self.L = 14, 8, 7
        self.start_board(chess960=chess960)

    def start_board(self, chess960=False):
        if chess960:
            print("Starting chess960")
            self._start_id = random.randint(0, 959)
        else:
            self._start_id = -1
        self._indices = []            # policy indices played so far
        self._moves_played = []
        self._cache = None
        self._frames = []             # per-position absolute piece planes + repetition flags (for .representation)
        self.board = chess.Board(self)
        self._record_frame()

    # -- engine plumbing -----------------------------------------------------------------------------
    def _engine(self):
        eng = runtime.get_engine()
        owner = getattr(eng, "owner", None)
        n = len(self._indices)
        if owner is not None and owner[0] is self and owner[1] <= n and eng.n_games == 1:
            for idx in self._indices[owner[1]:]:
                eng.push([0], [idx])
        else:
            eng.reset([self._start_id])
            for idx in self._indices:
                eng.push([0], [idx])
        eng.owner = (self, n)
        return eng

    def _state(self):
        if self._cache is None:
            eng = self._engine()
            pos = eng.positions()[0]
            idx, cnt = eng.legal_moves()
            white = bool(pos.turn)
            qp = {}
            moves = []
            for i in idx[0, :cnt[0]]:
                m = index_move(int(i), white)
                if m.promotion is None and ((pos.pieces[0 if white else 6] >> m.from_square) & 1) and (m.to_square // 8 in (0, 7)):
                    m = chess.make_move(m.from_square, m.to_square, chess.QUEEN)
                moves.append(m)
            self._cache = {"pos": pos, "moves": moves, "indices": [int(i) for i in idx[0, :cnt[0]]]}
        return self._cache

    def _push(self, move, check=True):
        st = self._state()
        idx = move_index(move, bool(st["pos"].turn))
        if check and idx not in st["indices"]:
            raise ValueError("Invalid move")
        eng = self._engine()
        try:
            eng.push([0], [idx])
        except ValueError:
            eng.owner = None
            raise
        self._indices.append(idx)
        self._moves_played.append(move)
        eng.owner = (self, len(self._indices))
        self._cache = None
        self._record_frame()

    def _record_frame(self):
        pos = self._state()["pos"]
        pieces = np.zeros((12, 8, 8), dtype=bool)
        for i in range(12):
            bits = np.unpackbits(np.array([pos.pieces[i]], dtype="<u8").view(np.uint8), bitorder="little")
            pieces[i] = bits.reshape(8, 8)
        self._frames.insert(0, (pieces, bool(pos.rep_flags & 1), bool(pos.rep_flags & 2)))
        del self._frames[self.T:]

    # -- reference API -------------------------------------------------------------------------------
    def move_piece(self, move):
        self._push(move, check=True)

    def get_representation(self):
        """bool[119,8,8] in the side-to-move's perspective, produced by the CUDA encoder (szb_encode)"""
        planes, _ = self._engine().encode(want_mask=False)
        return torch.from_numpy(runtime.unpack_planes(planes[0]))

    def _stack(self, white_first):
        """absolute-orientation plane stack like the reference's .representation / .black_representation"""
        out = np.zeros((119, 8, 8), dtype=bool)
        for t, (pieces, r2, r3) in enumerate(self._frames):
            out[14 * t:14 * t + 12] = pieces if white_first else np.concatenate([pieces[6:], pieces[:6]])
            out[14 * t + 12], out[14 * t + 13] = r2, r3
        b = self.board
        started = len(self._indices) > 0
        wk = b.has_kingside_castling_rights(True) if started else True
        wq = b.has_queenside_castling_rights(True) if started else True
        bk = b.has_kingside_castling_rights(False) if started else True
        bq = b.has_queenside_castling_rights(False) if started else True
        out[112] = white_first
        out[113] = started
        out[114:118] = np.array([wk, wq, bk, bq] if white_first else [bk, bq, wk, wq])[:, None, None]
        out[118] = started and b.halfmove_clock > 0
        return torch.from_numpy(out)

    @property
    def representation(self):
        return self._stack(True)

    @property
    def black_representation(self):
        return self._stack(False)

    def get_moves(self):
        return list(self.board.legal_moves)

    def get_initial_state(self):
        return self.board

    def get_valid_moves(self, state):
        return list(state.legal_moves)

    def get_value_and_terminated(self):
        o = self.board.outcome()
        if o is None:
            return 0, False
        return (0 if o.winner is None else -1), True

    def get_opponent(self, player):
        return -player

    def get_opponent_value(self, value):
        return -value
