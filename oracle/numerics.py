"""Bit-level restatement of the fp32 arithmetic the reference search asks torch (CPU) to do.

TEST INFRASTRUCTURE ONLY (oracle).  Plain numpy, one IEEE operation at a time, so that the CUDA
kernels in sigma-zero_b200/csrc/tree.cuh can be checked without torch in the loop and so that the
torch build on the GPU box's host can itself be probed (tests/test_numerics.py).

  puct_scores      <- mctsnode.py:23-37  (Node.select / Node.get_ucb)
  cascade_sum      <- mcts.py:79         (torch.sum over the masked fp32[4672] policy)
  normalise_policy <- mcts.py:77-81
  noisy_prior      <- mcts.py:91-96      (Dirichlet over ONE category == constant)

PARITY STATUS: pinned against torch itself (the reference's arithmetic library) by
tests/test_numerics.py on random vectors and by tests/golden/numerics_*.npz generated with torch 2.11.
"""
import math

import numpy as np

F = np.float32


def puct_scores(child_n, child_w, child_p, parent_n, c):
    """UCB of every child, exactly as torch evaluates mctsnode.py:33-37 in fp32.

    child_n: int visit counts, child_w: float64 value sums, child_p: fp32-exact priors.
    """
    n = np.asarray(child_n, dtype=np.int64)
    wf = np.asarray(child_w, dtype=np.float64).astype(F)          # torch.tensor([python floats]) -> fp32 RNE
    p = np.asarray(child_p, dtype=np.float64).astype(F)
    d = n.astype(F) + F(1e-6)                                     # vc + 1e-6 (int64 -> fp32 promotion)
    q = F(1.0) - ((wf / d) + F(1.0)) / F(2.0)
    r = F(1.0) / (n + 1).astype(F)                                # Tensor.__rtruediv__ == reciprocal() * other
    u = r * F(math.sqrt(parent_n))
    return (q + ((F(c) * u) * p)).astype(F)


def puct_select(child_n, child_w, child_p, parent_n, c):
    """torch.argmax: first maximum wins."""
    return int(np.argmax(puct_scores(child_n, child_w, child_p, parent_n, c)))


def cascade_sum(x):
    """ATen's contiguous fp32 sum (cascade_sum, 8 lanes x 4 ILP rows, 16-step levels) for len(x) % 32 == 0.

    element e -> vector v = e // 8, lane = e % 8, row k = v % 4, step i = v // 4.
    """
    x = np.asarray(x, dtype=F)
    n = x.shape[0]
    assert n % 32 == 0, "restated for whole 32-element steps only (4672 = 146 * 32)"
    steps = n // 32
    level_power = max(4, (int(steps - 1).bit_length() if steps > 1 else 0) // 4)
    level_step = 1 << level_power
    level_mask = level_step - 1
    xs = x.reshape(steps, 32)                        # [step, row*8 + lane]
    acc = np.zeros((4, 32), dtype=F)                 # 4 cascade levels
    i = 0
    while i + level_step <= steps:
        for j in range(level_step):
            acc[0] = acc[0] + xs[i + j]
        i += level_step
        for lvl in range(1, 4):
            acc[lvl] = acc[lvl] + acc[lvl - 1]
            acc[lvl - 1] = 0
            mask = level_mask << (lvl * level_power)
            if (i & mask) != 0:
                break
    while i < steps:
        acc[0] = acc[0] + xs[i]
        i += 1
    for lvl in range(1, 4):
        acc[0] = acc[0] + acc[lvl]
    rows = acc[0].reshape(4, 8)
    part = rows[0].copy()
    for k in range(1, 4):
        part = part + rows[k]
    total = F(0.0)
    for lane in range(8):
        total = F(total + part[lane])
    return F(total)


def normalise_policy(softmax_all, mask):
    """mcts.py:77-79: p = softmax * mask; p /= torch.sum(p).  Both fp32[4672]."""
    p = (np.asarray(softmax_all, dtype=F) * np.asarray(mask, dtype=F)).astype(F)
    return (p / cascade_sum(p)).astype(F)


NOISE_CONST = F(0.25) * np.nextafter(F(1.0), F(0.0))        # eps * Dirichlet([0.3]).sample() == 0.25 * 0.99999994


def noisy_prior(p):
    """mcts.py:91-96 with learning=True: (1-eps)*p + eps*noise where the one-category Dirichlet sample is
    the constant nextafter(1, 0)."""
    p = np.asarray(p, dtype=F)
    return (F(0.75) * p + NOISE_CONST).astype(F)
