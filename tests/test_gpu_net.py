"""GPU parity: the CUDA policy/value network against the reference architecture on CPU fp32 (torch).
Tolerances are the north-star's: <= 1e-5 abs for the fp32 path, <= 2e-2 abs for the bf16 tcgen05 path, on the
post-softmax policy and the tanh value."""
import os

import numpy as np
import pytest
import torch

from oracle import hash_eval, ref_path
from tests import util

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2


def _positions(n, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for g in range(n):
        c960 = g % 2 == 1
        og = util.oracle_game(c960, int(rng.integers(960)) if c960 else 518)
        for _ in range(int(rng.integers(0, 60))):
            legal = list(og.board.legal_moves)
            if og.board.outcome() is not None or not legal:
                break
            og.move_piece(legal[rng.integers(len(legal))])
        out.append(og.get_representation())
    return np.stack(out)


def _randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
            m.running_mean = 0.1 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var = 0.5 + torch.rand(m.running_var.shape, generator=g)


@pytest.fixture(scope="module")
def setup():
    from sigma_zero_b200.engine import Engine
    torch.manual_seed(0)
    model = ref_path.build_policy_nn().eval()
    eng = Engine(max_games=64, max_searches=16)
    eng.load_state_dict(model.state_dict())
    yield eng, model
    eng.close()


def _torch_forward(model, planes_bool):
    with torch.no_grad():
        p, v = model(torch.from_numpy(planes_bool.astype(np.float32)), inference=True)
    return p.numpy(), v.numpy().ravel()


def test_fp32_path_matches_torch_cpu(setup):
    from sigma_zero_b200.engine import EVAL_NET_FP32
    eng, model = setup
    x = _positions(6)
    packed = np.stack([hash_eval.pack_planes(p) for p in x])
    pol, val = eng.net_forward(packed, EVAL_NET_FP32)
    rp, rv = _torch_forward(model, x)
    assert np.abs(pol - rp).max() <= FP32_TOL, np.abs(pol - rp).max()
    assert np.abs(val - rv).max() <= FP32_TOL, np.abs(val - rv).max()
    assert np.abs(pol.sum(1) - 1).max() < 1e-4


def test_fp32_path_matches_golden(setup, golden_dir):
    from sigma_zero_b200.engine import EVAL_NET_FP32
    eng, _ = setup
    z = np.load(os.path.join(golden_dir, "network.npz"))
    pol, val = eng.net_forward(z["planes"], EVAL_NET_FP32)
    assert np.abs(pol - z["policy"]).max() <= FP32_TOL
    assert np.abs(val - z["value"].ravel()).max() <= FP32_TOL


def test_bf16_tcgen05_path_within_tolerance(setup):
    from sigma_zero_b200.engine import EVAL_NET_BF16
    eng, model = setup
    x = _positions(10, seed=3)
    packed = np.stack([hash_eval.pack_planes(p) for p in x])
    pol, val = eng.net_forward(packed, EVAL_NET_BF16)
    rp, rv = _torch_forward(model, x)
    assert np.isfinite(pol).all() and np.isfinite(val).all()
    assert np.abs(pol - rp).max() <= BF16_TOL, np.abs(pol - rp).max()
    assert np.abs(val - rv).max() <= BF16_TOL, np.abs(val - rv).max()
    # the absolute bound alone is vacuous on this head (every prior of a freshly initialised network is ~2e-4): also bound the
    # error relative to each board's largest prior, and require a normalised, non-degenerate policy
    rel = (np.abs(pol - rp).max(axis=1) / rp.max(axis=1)).max()
    assert rel <= 5e-2, rel
    assert np.abs(pol.sum(axis=1) - 1).max() < 1e-3 and pol.max() > 1e-4


def test_logits_and_folded_batchnorm(setup):
    """non-trivial BatchNorm statistics exercise the folding; logits (inference=False) compared directly"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, EVAL_NET_FP32, Engine
    torch.manual_seed(1)
    model = ref_path.build_policy_nn().eval()
    _randomise_bn(model, 5)
    eng = Engine(max_games=8, max_searches=4)
    eng.load_state_dict(model.state_dict())
    x = _positions(5, seed=9)
    packed = np.stack([hash_eval.pack_planes(p) for p in x])
    with torch.no_grad():
        rl, rv = model(torch.from_numpy(x.astype(np.float32)), inference=False)
    rl, rv = rl.numpy(), rv.numpy().ravel()
    scale = max(1.0, float(np.abs(rl).max()))
    l32, v32 = eng.net_forward(packed, EVAL_NET_FP32, logits=True)
    assert np.abs(l32 - rl).max() <= 2e-5 * scale, (np.abs(l32 - rl).max(), scale)
    assert np.abs(v32 - rv).max() <= FP32_TOL
    l16, v16 = eng.net_forward(packed, EVAL_NET_BF16, logits=True)
    assert np.abs(l16 - rl).max() <= 3e-2 * scale, (np.abs(l16 - rl).max(), scale)
    assert np.abs(v16 - rv).max() <= BF16_TOL
    eng.close()


@pytest.mark.parametrize("evaluator", [0, 1])
def test_batch_invariance(setup, evaluator):
    """a position's outputs must not depend on its batch slot or the batch size (reproducible visit counts)"""
    eng, _ = setup
    x = _positions(7, seed=4)
    packed = np.stack([hash_eval.pack_planes(p) for p in x])
    pol_a, val_a = eng.net_forward(packed, evaluator)
    perm = np.array([3, 0, 6, 2, 5, 1, 4])
    big = np.concatenate([packed[perm], packed, packed[:5]])
    pol_b, val_b = eng.net_forward(big, evaluator)
    assert np.array_equal(pol_b[:7], pol_a[perm]) and np.array_equal(val_b[:7], val_a[perm])
    assert np.array_equal(pol_b[7:14], pol_a) and np.array_equal(val_b[14:], val_a[:5])
    pol_c, val_c = eng.net_forward(packed[2:3], evaluator)
    assert np.array_equal(pol_c[0], pol_a[2]) and val_c[0] == val_a[2]


def test_tower_kernel_variants_bit_identical(monkeypatch):
    """the three bf16 tower implementations (single-CTA per layer, CTA-pair per layer, CTA-pair one persistent launch
    with per-item dependency counters) run the same MMAs in the same K order: their outputs must be bit-identical,
    at a batch that spans several waves of the persistent kernel and is not a multiple of the 4-board pair tile"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine
    torch.manual_seed(2)
    model = ref_path.build_policy_nn().eval()
    _randomise_bn(model, 11)
    x = _positions(9, seed=21)
    packed = np.stack([hash_eval.pack_planes(p) for p in x])
    n = 4 * 74 * 2 + 3                      # > 2 items per CTA pair per layer, ragged last tile
    big = packed[np.arange(n) % len(packed)]
    outs = []
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("SZB_TOWER_MODE", mode)
        eng = Engine(max_games=n, max_searches=4)
        eng.load_state_dict(model.state_dict())
        outs.append(eng.net_forward(big, EVAL_NET_BF16, logits=True))
        eng.close()
    for l, v in outs[1:]:
        assert np.array_equal(l, outs[0][0]) and np.array_equal(v, outs[0][1])
    # and the batch really is periodic: every copy of a position gives the same logits
    assert np.array_equal(outs[2][0][:len(packed)], outs[2][0][len(packed):2 * len(packed)])


def test_tower_n_split_bit_identical(monkeypatch):
    """small batches cut every (layer, tile) of the persistent tower launch into 2, 4 or 8 work items of N / nsplit output
    channels (more CTA pairs per layer, shorter dependency chain): the same MMAs per output in the same K order, so forced
    splits 1 / 2 / 4 / 8 and the per-layer launch must agree bit for bit -- ragged last tile, residual and policy layers included"""
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine
    torch.manual_seed(3)
    model = ref_path.build_policy_nn().eval()
    _randomise_bn(model, 13)
    x = _positions(9, seed=22)
    packed = np.stack([hash_eval.pack_planes(p) for p in x])
    for n in (1, 37, 150):
        big = packed[np.arange(n) % len(packed)]
        outs = []
        for mode, split in (("1", None), ("2", "1"), ("2", "2"), ("2", "4"), ("2", "8"), ("2", None)):
            monkeypatch.setenv("SZB_TOWER_MODE", mode)
            if split is None:
                monkeypatch.delenv("SZB_TOWER_NSPLIT", raising=False)
            else:
                monkeypatch.setenv("SZB_TOWER_NSPLIT", split)
            eng = Engine(max_games=n, max_searches=4)
            eng.load_state_dict(model.state_dict())
            outs.append(eng.net_forward(big, EVAL_NET_BF16, logits=True))
            eng.close()
        for l, v in outs[1:]:
            assert np.array_equal(l, outs[0][0]) and np.array_equal(v, outs[0][1]), n
