"""Clean-room stand-in for the subset of python-chess 1.10.0 that the reference hot path uses.

TEST INFRASTRUCTURE ONLY (oracle).  python-chess (`chess==1.10.0`, reference environment.yml:21) is a
third-party dependency that is absent from /root/reference and cannot be installed offline.  This module
lets the *unmodified* reference files (mcts.py, mctsnode.py, chess_tensor.py, sim.py) run by providing a
module named `chess` with exactly the names they touch (call sites: SURVEY.md §2.2):

    WHITE/BLACK, PAWN..KING, SQUARES, Color, Piece, Move(.from_square,.to_square,.promotion,.uci(),
    .from_uci), Board(), Board.from_chess960_pos(n), .turn, .move_stack, .halfmove_clock, .legal_moves,
    .piece_at, .push, .pop, .is_repetition, .has_kingside/queenside_castling_rights, .is_game_over,
    .outcome().winner, .result(), deepcopy, str().

The rules themselves live in oracle/chess_oracle.c (mailbox engine); this file is only the object model.
PARITY STATUS: "parity unpinned" by the reference (no tests / golden vectors at this boundary); pinned
by public perft tables, Scharnagl anchors and hand-derived vectors in tests/.
"""
from __future__ import annotations

import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
import importlib.util as _ilu

_spec = _ilu.spec_from_file_location("_szb_oracle_build", os.path.join(os.path.dirname(_here), "build.py"))
_oracle_build = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_oracle_build)

_L = ctypes.CDLL(_oracle_build.build())
_vp, _i, _u16 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint16
for _name, _res, _args in [
    ("ora_game_new_startpos", _vp, [_i]),
    ("ora_game_new_fen", _vp, [ctypes.c_char_p, _i]),
    ("ora_game_copy", _vp, [_vp]),
    ("ora_game_free", None, [_vp]),
    ("ora_turn", _i, [_vp]),
    ("ora_ply", _i, [_vp]),
    ("ora_halfmove", _i, [_vp]),
    ("ora_ep_square", _i, [_vp]),
    ("ora_is_chess960", _i, [_vp]),
    ("ora_piece_at", _i, [_vp, _i]),
    ("ora_rights", _i, [_vp, _i]),
    ("ora_board", None, [_vp, ctypes.c_void_p]),
    ("ora_legal_moves", _i, [_vp, ctypes.c_void_p]),
    ("ora_is_legal", _i, [_vp, _u16]),
    ("ora_move_at", _i, [_vp, _i]),
    ("ora_push", None, [_vp, _u16]),
    ("ora_pop", _i, [_vp]),
    ("ora_is_check", _i, [_vp]),
    ("ora_has_castling", _i, [_vp, _i, _i]),
    ("ora_is_repetition", _i, [_vp, _i]),
    ("ora_outcome", _i, [_vp]),
    ("ora_perft", ctypes.c_uint64, [_vp, _i]),
    ("ora_divide", _i, [_vp, _i, ctypes.c_void_p, ctypes.c_void_p]),
    ("ora_chess960_backrank", None, [_i, ctypes.c_char_p]),
    ("ora_export", None, [_vp] + [ctypes.c_void_p] * 7),
]:
    _f = getattr(_L, _name)
    _f.restype = _res
    _f.argtypes = _args

Color = bool
WHITE = True
BLACK = False
PieceType = int
PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING = range(1, 7)
PIECE_SYMBOLS = [None, "p", "n", "b", "r", "q", "k"]
Square = int
SQUARES = list(range(64))
(A1, B1, C1, D1, E1, F1, G1, H1) = range(8)
(A8, B8, C8, D8, E8, F8, G8, H8) = range(56, 64)
FILE_NAMES = "abcdefgh"
RANK_NAMES = "12345678"
SQUARE_NAMES = [f + r for r in RANK_NAMES for f in FILE_NAMES]
STARTING_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"


def square_file(sq: int) -> int:
    return sq & 7


def square_rank(sq: int) -> int:
    return sq >> 3


def square_name(sq: int) -> str:
    return SQUARE_NAMES[sq]


def parse_square(name: str) -> int:
    return SQUARE_NAMES.index(name)


class Piece:
    __slots__ = ("piece_type", "color")

    def __init__(self, piece_type: int, color: bool):
        self.piece_type = piece_type
        self.color = color

    def symbol(self) -> str:
        s = PIECE_SYMBOLS[self.piece_type]
        return s.upper() if self.color else s

    def __eq__(self, other):
        return isinstance(other, Piece) and (self.piece_type, self.color) == (other.piece_type, other.color)

    def __hash__(self):
        return hash((self.piece_type, self.color))

    def __repr__(self):
        return "Piece.from_symbol(%r)" % self.symbol()


class Move:
    __slots__ = ("from_square", "to_square", "promotion", "drop")

    def __init__(self, from_square: int, to_square: int, promotion=None, drop=None):
        self.from_square = from_square
        self.to_square = to_square
        self.promotion = promotion
        self.drop = drop

    def uci(self) -> str:
        if self.from_square == self.to_square == 0 and not self.promotion:
            return "0000"
        s = SQUARE_NAMES[self.from_square] + SQUARE_NAMES[self.to_square]
        if self.promotion:
            s += PIECE_SYMBOLS[self.promotion]
        return s

    @classmethod
    def from_uci(cls, uci: str) -> "Move":
        if uci == "0000":
            return cls(0, 0)
        if len(uci) not in (4, 5):
            raise ValueError("expected uci string to be of length 4 or 5: %r" % uci)
        try:
            f = SQUARE_NAMES.index(uci[0:2])
            t = SQUARE_NAMES.index(uci[2:4])
            promo = PIECE_SYMBOLS.index(uci[4]) if len(uci) == 5 else None
        except ValueError:
            raise ValueError("invalid uci: %r" % uci)
        if f == t:
            raise ValueError("invalid uci (use 0000 for null moves): %r" % uci)
        return cls(f, t, promo)

    def _code(self) -> int:
        return self.from_square | (self.to_square << 6) | ((self.promotion or 0) << 12)

    @classmethod
    def _from_code(cls, code: int) -> "Move":
        p = (code >> 12) & 7
        return cls(code & 63, (code >> 6) & 63, p or None)

    def __bool__(self):
        return bool(self.from_square or self.to_square or self.promotion)

    def __eq__(self, other):
        return isinstance(other, Move) and self._code() == other._code()

    def __hash__(self):
        return hash(self._code())

    def __repr__(self):
        return "Move.from_uci(%r)" % self.uci()

    def __str__(self):
        return self.uci()


class Outcome:
    def __init__(self, termination: int, winner):
        self.termination = termination
        self.winner = winner

    def result(self) -> str:
        return "1/2-1/2" if self.winner is None else ("1-0" if self.winner else "0-1")


class _MoveStack:
    """Read-only view of the played moves (len(), indexing, iteration, truthiness)."""

    def __init__(self, board: "Board"):
        self._b = board

    def __len__(self):
        return _L.ora_ply(self._b._g)

    def __getitem__(self, i):
        n = len(self)
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(n))]
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return Move._from_code(_L.ora_move_at(self._b._g, i))

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class LegalMoveGenerator:
    def __init__(self, board: "Board"):
        self._b = board

    def _codes(self):
        buf = (ctypes.c_uint16 * 512)()
        n = _L.ora_legal_moves(self._b._g, buf)
        return buf[:n]

    def __iter__(self):
        return iter([Move._from_code(c) for c in self._codes()])

    def __len__(self):
        return len(self._codes())

    def count(self):
        return len(self)

    def __bool__(self):
        return len(self) > 0

    def __contains__(self, move: Move):
        return bool(_L.ora_is_legal(self._b._g, move._code()))


class Board:
    def __init__(self, fen: str | None = STARTING_FEN, *, chess960: bool = False):
        self.chess960 = chess960
        if fen == STARTING_FEN and not chess960:
            self._g = _L.ora_game_new_startpos(-1)
        else:
            self._g = _L.ora_game_new_fen(fen.encode(), int(chess960))

    @classmethod
    def from_chess960_pos(cls, scharnagl: int) -> "Board":
        if not 0 <= scharnagl <= 959:
            raise ValueError("chess960 position index not 0 <= %r <= 959" % scharnagl)
        b = cls.__new__(cls)
        b.chess960 = True
        b._g = _L.ora_game_new_startpos(scharnagl)
        return b

    def __del__(self):
        g, self._g = getattr(self, "_g", None), None
        if g and _L is not None:
            _L.ora_game_free(g)

    def copy(self) -> "Board":
        b = Board.__new__(Board)
        b.chess960 = self.chess960
        b._g = _L.ora_game_copy(self._g)
        return b

    def __copy__(self):
        return self.copy()

    def __deepcopy__(self, memo):
        return self.copy()

    # -- state ---------------------------------------------------------------------------------
    @property
    def turn(self) -> bool:
        return bool(_L.ora_turn(self._g))

    @property
    def halfmove_clock(self) -> int:
        return _L.ora_halfmove(self._g)

    @property
    def ep_square(self):
        e = _L.ora_ep_square(self._g)
        return None if e < 0 else e

    @property
    def fullmove_number(self) -> int:
        return 1 + (_L.ora_ply(self._g) + (0 if self._start_turn() else 1)) // 2

    def _start_turn(self) -> bool:
        return bool(_L.ora_turn(self._g)) == (_L.ora_ply(self._g) % 2 == 0)

    @property
    def move_stack(self):
        return _MoveStack(self)

    @property
    def legal_moves(self) -> LegalMoveGenerator:
        return LegalMoveGenerator(self)

    def piece_at(self, square: int):
        c = _L.ora_piece_at(self._g, square)
        if c == 0:
            return None
        return Piece(abs(c), c > 0)

    def piece_type_at(self, square: int):
        c = _L.ora_piece_at(self._g, square)
        return abs(c) or None

    def is_legal(self, move: Move) -> bool:
        return move in self.legal_moves

    # -- moves ---------------------------------------------------------------------------------
    def push(self, move: Move) -> None:
        _L.ora_push(self._g, move._code())

    def pop(self) -> Move:
        m = self.move_stack[-1]
        _L.ora_pop(self._g)
        return m

    def push_uci(self, uci: str) -> Move:
        m = Move.from_uci(uci)
        if m not in self.legal_moves:
            raise ValueError("illegal uci: %r in %s" % (uci, self.fen()))
        self.push(m)
        return m

    # -- queries -------------------------------------------------------------------------------
    def is_check(self) -> bool:
        return bool(_L.ora_is_check(self._g))

    def is_repetition(self, count: int = 3) -> bool:
        return bool(_L.ora_is_repetition(self._g, count))

    def has_kingside_castling_rights(self, color: bool) -> bool:
        return bool(_L.ora_has_castling(self._g, int(color), 1))

    def has_queenside_castling_rights(self, color: bool) -> bool:
        return bool(_L.ora_has_castling(self._g, int(color), 0))

    def outcome(self, *, claim_draw: bool = False):
        if claim_draw:
            raise NotImplementedError("claim_draw is not used by the reference path")
        o = _L.ora_outcome(self._g)
        if o == 0:
            return None
        return Outcome(o, (not self.turn) if o == 1 else None)

    def is_game_over(self, *, claim_draw: bool = False) -> bool:
        return self.outcome(claim_draw=claim_draw) is not None

    def is_checkmate(self) -> bool:
        return _L.ora_outcome(self._g) == 1

    def result(self, *, claim_draw: bool = False) -> str:
        o = self.outcome(claim_draw=claim_draw)
        return o.result() if o else "*"

    def perft(self, depth: int) -> int:
        return int(_L.ora_perft(self._g, depth))

    def divide(self, depth: int) -> dict:
        mv = (ctypes.c_uint16 * 512)()
        nd = (ctypes.c_uint64 * 512)()
        n = _L.ora_divide(self._g, depth, mv, nd)
        return {Move._from_code(mv[i]).uci(): int(nd[i]) for i in range(n)}

    # -- text ----------------------------------------------------------------------------------
    def _symbols(self):
        buf = (ctypes.c_int8 * 64)()
        _L.ora_board(self._g, buf)
        return [None if c == 0 else (PIECE_SYMBOLS[abs(c)].upper() if c > 0 else PIECE_SYMBOLS[abs(c)]) for c in buf]

    def __str__(self) -> str:
        s = self._symbols()
        return "\n".join(" ".join(s[r * 8 + f] or "." for f in range(8)) for r in range(7, -1, -1))

    def board_fen(self) -> str:
        s = self._symbols()
        rows = []
        for r in range(7, -1, -1):
            row, empty = "", 0
            for f in range(8):
                c = s[r * 8 + f]
                if c is None:
                    empty += 1
                else:
                    row += (str(empty) if empty else "") + c
                    empty = 0
            rows.append(row + (str(empty) if empty else ""))
        return "/".join(rows)

    def fen(self) -> str:
        """Shredder-style FEN (castling rights as rook files) -- unambiguous for Chess960."""
        rights = ""
        for color, up in ((1, True), (0, False)):
            m = _L.ora_rights(self._g, color)
            for f in range(7, -1, -1):
                if m & (1 << f):
                    rights += FILE_NAMES[f].upper() if up else FILE_NAMES[f]
        ep = self.ep_square
        return "%s %s %s %s %d %d" % (
            self.board_fen(), "w" if self.turn else "b", rights or "-",
            SQUARE_NAMES[ep] if ep is not None else "-", self.halfmove_clock, self.fullmove_number)

    def _export(self):
        """(bb12, turn, rights_w, rights_b, ep, halfmove, ply) in the product's szb_pos wire layout."""
        bb = (ctypes.c_uint64 * 12)()
        v = [ctypes.c_int() for _ in range(6)]
        _L.ora_export(self._g, bb, *[ctypes.byref(x) for x in v])
        return [int(x) for x in bb], *[x.value for x in v]

    def __repr__(self):
        return "Board(%r, chess960=%r)" % (self.fen(), self.chess960)


def chess960_backrank(scharnagl: int) -> str:
    buf = ctypes.create_string_buffer(8)
    _L.ora_chess960_backrank(scharnagl, buf)
    return buf.raw.decode()
