"""Host logic of the fine-tuning path (no GPU): the batch schedule train_on_records / train_on_records_cuda share, and the routing between
the library's trainer and the torch oracle."""
import numpy as np
import pytest


def test_batch_schedule_matches_dataloader_drop_last():
    """a fresh permutation per epoch; the short last batch is dropped (DataLoader(drop_last=True), train_RL.py:241-246) unless the whole
    record set is smaller than one batch; BatchNorm needs more than one sample; the schedule is a function of the seed alone"""
    from sigma_zero_b200.train_RL import _batches
    b = list(_batches(10, 4, 2, seed=3))
    assert [len(r) for r in b] == [4, 4, 4, 4]                       # 10 = 4 + 4 + (2 dropped), twice
    assert sorted(np.concatenate(b[:2]).tolist()) != sorted(np.concatenate(b[2:]).tolist()) or not np.array_equal(b[0], b[2])
    for epoch in (b[:2], b[2:]):
        rows = np.concatenate(epoch)
        assert len(set(rows.tolist())) == 8 and rows.min() >= 0 and rows.max() < 10
    again = list(_batches(10, 4, 2, seed=3))
    assert all(np.array_equal(x, y) for x, y in zip(b, again))
    assert not all(np.array_equal(x, y) for x, y in zip(b, _batches(10, 4, 2, seed=4)))
    assert [len(r) for r in _batches(3, 8, 1, seed=0)] == [3]         # fewer records than one batch: one short batch
    assert list(_batches(1, 8, 1, seed=0)) == []                      # a single sample cannot be batch-normalised
    assert [len(r) for r in _batches(9, 4, 1, seed=0)] == [4, 4]      # the trailing single record is dropped


def test_backend_routing_without_a_gpu():
    """device="cpu" (or a CPU model with no device given) takes the torch path; asking for the CUDA trainer without a GPU fails loudly"""
    import torch
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.network import policyNN
    from sigma_zero_b200.train_RL import train_on_records
    if torch.cuda.is_available():
        pytest.skip("routing without a GPU is what this checks")
    rng = np.random.default_rng(0)
    n = 4
    rec = {"states": rng.integers(0, 2 ** 63, size=(n, 119), dtype=np.int64).astype(np.uint64), "pi_index": np.array([1, 7, 9, 40], np.uint16),
           "pi_prob": np.ones(4, np.float32), "pi_off": np.arange(n + 1, dtype=np.int64), "z": np.array([1, -1, 0, 1], np.int8)}
    torch.manual_seed(0)
    model = policyNN({})
    hist = train_on_records(model, rec, epochs=1, batch_size=4)       # CPU parameters, no device named -> torch autograd
    assert len(hist) == 1 and np.isfinite(hist[0]).all() and not model.training
    with pytest.raises(_lib.SzbError):
        train_on_records(model, rec, epochs=1, batch_size=4, backend="cuda")
