"""Minimal stand-in for the `chess` names the reference's hot path touches, served by the GPU engine.

The reference takes its rules from python-chess (environment.yml:21).  When python-chess is importable the
facade hands out real `chess.Move` objects; otherwise these light-weight equivalents are used.  No chess rule is
evaluated here: legality, outcomes, repetition and castling rights all come from libszb200 (szb_games_get,
szb_legal_moves).  Host logic only."""
import numpy as np

WHITE, BLACK = True, False
Color = bool
PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING = range(1, 7)
PIECE_SYMBOLS = [None, "p", "n", "b", "r", "q", "k"]
SQUARES = list(range(64))
SQUARE_NAMES = [f + r for r in "12345678" for f in "abcdefgh"]

try:                                      # prefer the real library when the user has it
    import importlib

    _real = importlib.import_module("chess")
    if not hasattr(_real, "Board") or getattr(_real, "__file__", "").endswith("chess_compat.py") or "oracle" in getattr(_real, "__file__", ""):
        _real = None
except Exception:                         # pragma: no cover
    _real = None


class Move:
    __slots__ = ("from_square", "to_square", "promotion")

    def __init__(self, from_square, to_square, promotion=None):
        self.from_square, self.to_square, self.promotion = from_square, to_square, promotion

    def uci(self):
        return SQUARE_NAMES[self.from_square] + SQUARE_NAMES[self.to_square] + (PIECE_SYMBOLS[self.promotion] if self.promotion else "")

    @classmethod
    def from_uci(cls, uci):
        if len(uci) not in (4, 5):
            raise ValueError("expected uci string to be of length 4 or 5: %r" % uci)
        return cls(SQUARE_NAMES.index(uci[:2]), SQUARE_NAMES.index(uci[2:4]), PIECE_SYMBOLS.index(uci[4]) if len(uci) == 5 else None)

    def __eq__(self, other):
        return (self.from_square, self.to_square, self.promotion or None) == \
               (getattr(other, "from_square", None), getattr(other, "to_square", None), getattr(other, "promotion", None) or None)

    def __hash__(self):
        return hash((self.from_square, self.to_square, self.promotion or 0))

    def __repr__(self):
        return "Move.from_uci(%r)" % self.uci()

    __str__ = uci


def make_move(from_square, to_square, promotion=None):
    if _real is not None:
        return _real.Move(from_square, to_square, promotion)
    return Move(from_square, to_square, promotion)


class Piece:
    def __init__(self, piece_type, color):
        self.piece_type, self.color = piece_type, color

    def symbol(self):
        s = PIECE_SYMBOLS[self.piece_type]
        return s.upper() if self.color else s


class Outcome:
    def __init__(self, termination, winner):
        self.termination, self.winner = termination, winner

    def result(self):
        return "1/2-1/2" if self.winner is None else ("1-0" if self.winner else "0-1")


class Board:
    """View of one game that lives on the GPU.  `owner` is the ChessTensor that replays it into the engine."""

    def __init__(self, owner):
        self._owner = owner

    def _pos(self):
        return self._owner._state()["pos"]

    @property
    def turn(self):
        return bool(self._pos().turn)

    @property
    def chess960(self):
        return bool(self._pos().chess960)

    @property
    def halfmove_clock(self):
        return int(self._pos().halfmove_clock)

    @property
    def move_stack(self):
        return list(self._owner._moves_played)

    @property
    def legal_moves(self):
        return self._owner._state()["moves"]

    def piece_at(self, square):
        p = self._pos()
        for i in range(12):
            if (p.pieces[i] >> square) & 1:
                return Piece(i % 6 + 1, i < 6)
        return None

    def push(self, move):
        self._owner._push(move, check=False)

    def is_repetition(self, count=3):
        if count not in (2, 3):
            raise NotImplementedError("the engine tracks is_repetition(2) and (3), the ones the input planes need")
        return bool(self._pos().rep_flags & (1 if count == 2 else 2))

    def _castling(self, color, kingside):
        p = self._pos()
        rights = p.castling_w if color else p.castling_b
        king = p.pieces[5 if color else 11] & (0xFF if color else 0xFF << 56)
        if not king or not rights:
            return False
        kf = (int(king).bit_length() - 1) & 7
        return bool(rights >> (kf + 1)) if kingside else bool(rights & ((1 << kf) - 1))

    def has_kingside_castling_rights(self, color):
        return self._castling(color, True)

    def has_queenside_castling_rights(self, color):
        return self._castling(color, False)

    def outcome(self, claim_draw=False):
        o = int(self._pos().outcome)
        if o == 0:
            return None
        return Outcome(o, (not self.turn) if o == 1 else None)

    def is_game_over(self, claim_draw=False):
        return self._pos().outcome != 0

    def result(self, claim_draw=False):
        o = self.outcome()
        return o.result() if o else "*"

    def __str__(self):
        p = self._pos()
        rows = []
        for r in range(7, -1, -1):
            row = []
            for f in range(8):
                pc = self.piece_at(r * 8 + f)
                row.append(pc.symbol() if pc else ".")
            rows.append(" ".join(row))
        return "\n".join(rows)
