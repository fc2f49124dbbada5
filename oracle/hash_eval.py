"""Deterministic integer-hash evaluator (TEST INFRASTRUCTURE ONLY -- oracle side).

A stand-in "network" whose outputs are an exact function of the 119 input planes, computable bit-for-bit
both here (numpy) and on the GPU (sigma-zero_b200/csrc/evaluators.cuh: szb::hash_eval).  It lets the tree
kernels, the in-tree move generator and the plane encoder be checked for exact visit-count parity against
the reference search without any floating-point network in the loop, and because the hash covers every
plane of the 8-ply history, a single wrong plane bit anywhere in a search shows up as a diverged tree.

    words[k]  = sum over (row, col) of planes[k,row,col] << (row*8 + col)          k = 0..118
    h         = fold(words)                                    (splitmix-style 64-bit mixing)
    r_i       = mix(h + (i+1) * GOLDEN) >> 40                  24-bit integer per action i
    policy[i] = 0                      if r_i % 256 == 0       (exercises the zero-prior-drop rule)
              = u*u*u*u, u = (r_i + 1) * 2^-24 in fp32         otherwise (peaky, strictly positive)
    value     = (mix(h ^ VALUE_SALT) >> 40) * 2^-23 - 1        fp32 in [-1, 1)
"""
import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
VALUE_SALT = np.uint64(0xD6E8FEB86659FD93)
_IDX = (np.arange(4672, dtype=np.uint64) + np.uint64(1))


def _mix(z):
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * M1
        z = (z ^ (z >> np.uint64(27))) * M2
        return z ^ (z >> np.uint64(31))


def pack_planes(planes) -> np.ndarray:
    """bool[119,8,8] -> uint64[119], bit (row*8+col)."""
    b = np.asarray(planes).astype(bool).reshape(119, 64)
    weights = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    return (b.astype(np.uint64) * weights).sum(axis=1, dtype=np.uint64)


def fold(words) -> np.uint64:
    h = np.uint64(0x243F6A8885A308D3)
    with np.errstate(over="ignore"):
        for w in np.asarray(words, dtype=np.uint64):
            h = _mix((h ^ w) + GOLDEN)
    return h


def evaluate_words(words):
    h = fold(words)
    with np.errstate(over="ignore"):
        r = _mix(h + _IDX * GOLDEN) >> np.uint64(40)
        rv = _mix(h ^ VALUE_SALT) >> np.uint64(40)
    u = (r + np.uint64(1)).astype(np.float32) * np.float32(2.0 ** -24)
    p = ((u * u) * u) * u
    p = np.where((r % np.uint64(256)) == 0, np.float32(0.0), p).astype(np.float32)
    v = np.float32(np.float32(rv) * np.float32(2.0 ** -23) - np.float32(1.0))
    return p, v


def evaluator(planes):
    """search() evaluator signature: bool[119,8,8] -> (fp32[4672], fp32 value)."""
    return evaluate_words(pack_planes(planes))
