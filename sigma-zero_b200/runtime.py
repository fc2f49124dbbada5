"""Process-wide engine management for the drop-in facade: one libszb200 context per GPU, grown on demand,
with the model's weights pushed to it whenever they change.  Host logic only."""
import os

import numpy as np

from .engine import EVAL_NET_BF16, EVAL_NET_FP32, Engine

_engine = None
_weights_key = None

DEFAULT_PRECISION = os.environ.get("SZB_PRECISION", "bf16")     # "bf16" (tcgen05) or "fp32" (parity mode)


def device_ordinal():
    return int(os.environ.get("LOCAL_RANK", os.environ.get("SZB_DEVICE", "0")))


def get_engine(min_games=1, min_searches=800, leaves_per_tree=None):
    """An Engine with at least the requested capacity (re-created, and weights re-sent, when it must grow or when the
    search mode changes).  leaves_per_tree: None = keep the current engine's mode (1 for a new one)."""
    global _engine, _weights_key
    want_k = (_engine.leaves_per_tree if _engine else 1) if leaves_per_tree is None else int(leaves_per_tree)
    if _engine is None or _engine.max_games < min_games or _engine.max_searches < min_searches or _engine.leaves_per_tree != want_k:
        games = max(min_games, _engine.max_games if _engine else 1)
        searches = max(min_searches, _engine.max_searches if _engine else 1)
        if _engine is not None:
            _engine.close()
        _engine = Engine(max_games=games, max_searches=searches, device=device_ordinal(), leaves_per_tree=want_k)
        _engine.owner = None
        _weights_key = None
    return _engine


def leaves_of(args):
    """args['leaves_per_tree'] (an extension of the reference's args dict; default 1 = the reference's search)"""
    return int((args or {}).get('leaves_per_tree', 1))


def evaluator_of(model):
    prec = getattr(model, "precision", DEFAULT_PRECISION)
    return EVAL_NET_FP32 if prec == "fp32" else EVAL_NET_BF16


def sync_weights(engine, model):
    """Send the model's state_dict to the GPU library if it changed since the last call."""
    global _weights_key
    sd = model.state_dict()
    key = (id(model), id(engine), tuple(int(getattr(v, "_version", 0)) for v in sd.values()),
           tuple(v.data_ptr() for v in sd.values()))
    if key != _weights_key:
        engine.load_state_dict(sd)
        _weights_key = key


def unpack_planes(words):
    """uint64[..., 119] -> bool[..., 119, 8, 8] (host-side view change of the GPU encoder's output)"""
    w = np.ascontiguousarray(words, dtype="<u8")
    bits = np.unpackbits(w.view(np.uint8).reshape(w.shape + (8,)), axis=-1, bitorder="little")
    return bits.reshape(w.shape + (8, 8)).astype(bool)


def pack_planes(planes):
    """bool/float[..., 119, 8, 8] -> uint64[..., 119]"""
    p = np.asarray(planes) != 0
    flat = p.reshape(p.shape[:-2] + (64,))
    return np.packbits(flat, axis=-1, bitorder="little").view("<u8").reshape(p.shape[:-2])
