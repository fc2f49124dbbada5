"""GPU parity of the CUDA trainer (szb_train_*, csrc/train.cu) with the reference's fine-tuning step (train_RL.py:77-154) as torch
autograd computes it in fp32 on the same GPU: loss = mse_loss(v, z) + cross_entropy(logits, pi), Adam(1e-4, wd 1e-4), StepLR(500, 0.95).

Tolerances.  The trainer computes with bf16 tensor-core operands (fp32 accumulation, fp32 master weights): against fp32 autograd it
cannot be closer than bf16 allows through 40 BatchNorm'd layers, so the yardstick for forward values and gradients is torch's OWN
bf16-autocast run of the same step -- the trainer's error against fp32 may not exceed 1.25x autocast's error (+ 0.02) per tensor.
Everything that does not pass through bf16 is held tightly: the optimiser arithmetic (<= 1e-7 abs after a step), the policy-bias
gradient (softmax - pi, <= 1e-2 rel), the BatchNorm running statistics (<= 5e-3 rel), determinism and resume (bit-identical)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _records(n, seed=0):
    rng = np.random.default_rng(seed)
    bits = rng.random((n, 119, 64)) < 0.12
    states = np.packbits(bits, axis=-1, bitorder="little").view("<u8").reshape(n, 119).astype(np.uint64)
    idx, prob, off = [], [], [0]
    for _ in range(n):
        k = int(rng.integers(5, 45))
        ind = np.sort(rng.choice(4672, size=k, replace=False))
        p = rng.random(k).astype(np.float32) ** 3
        p /= p.sum()
        idx.extend(ind.tolist())
        prob.extend(p.tolist())
        off.append(len(idx))
    return {"states": states, "pi_index": np.array(idx, np.uint16), "pi_prob": np.array(prob, np.float32), "pi_off": np.array(off, np.int64),
            "z": rng.integers(-1, 2, n).astype(np.int8), "colour": np.ones(n, bool), "game": np.zeros(n, np.int32)}


def _model(seed=0):
    from sigma_zero_b200.network import policyNN
    torch.manual_seed(seed)
    model = policyNN({})
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(0.5 + torch.rand(mod.weight.shape, generator=g))
                mod.bias.copy_(0.2 * torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=g))
    return model


def _torch_batch(rec, rows, dev):
    from sigma_zero_b200 import records
    x = torch.from_numpy(records.unpack_states(rec, rows)).to(device=dev, dtype=torch.float32)
    pi = torch.from_numpy(records.dense_policy(rec, rows)).to(dev)
    z = torch.from_numpy(rec["z"][rows].astype(np.float32)).to(dev)
    return x, pi, z


def _torch_step(model, x, pi, z, autocast=False):
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        p, v = model.forward_torch(x)
    p, v = p.float(), v.float()
    mse = torch.nn.functional.mse_loss(v.squeeze(-1), z)
    ce = torch.nn.functional.cross_entropy(p, pi)
    (mse + ce).backward()
    return float(mse.detach()), float(ce.detach()), p.detach(), v.detach().squeeze(-1)


def _rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


@pytest.fixture(scope="module")
def setup():
    from sigma_zero_b200.engine import Engine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    eng = Engine(max_games=2, max_searches=8, device=0)
    yield eng, torch.device("cuda", 0)
    eng.close()


@pytest.mark.parametrize("batch", [32, 7])
def test_forward_and_every_gradient_against_autograd(setup, batch):
    """one step without update: losses, logits, value, all 129 gradient tensors, BatchNorm running buffers (odd batch: 7 boards)"""
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.trainer import Trainer
    eng, dev = setup
    rec = _records(64, seed=3)
    rows = (np.arange(batch, dtype=np.int32) * 5 + 1) % 64
    model = _model(0).to(dev).train()
    ref_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, pi, z = _torch_batch(rec, rows, dev)
    mse_t, ce_t, p_t, v_t = _torch_step(model, x, pi, z)
    grads_t = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    bn_t = {k: v.detach().clone() for k, v in model.state_dict().items() if "running" in k}
    model.load_state_dict(ref_sd)
    _, _, p_a, v_a = _torch_step(model, x, pi, z, autocast=True)
    grads_a = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    model.load_state_dict(ref_sd)
    tr = Trainer(eng, model, batch_size=32)
    try:
        tr.set_records(rec)
        mse_o, ce_o = tr.step(rows, flags=_lib.TRAIN_NO_UPDATE)
        assert abs(mse_o - mse_t) <= 5e-3 and abs(ce_o - ce_t) <= 5e-3, (mse_o, mse_t, ce_o, ce_t)
        lg, val = tr.activations(batch)
        assert _rel(lg, p_t) <= 1.25 * _rel(p_a, p_t) + 0.02
        assert (val - v_t.cpu()).abs().max().item() <= 2e-2 + 2 * (v_a - v_t).abs().max().item()
        g = tr.get_tensors(_lib.TRAIN_GRADS)
        assert set(g) == set(grads_t) and len(g) == 129
        worst = []
        for k in g:
            ro, ra = _rel(g[k], grads_t[k]), _rel(grads_a[k], grads_t[k])
            bound = 1.25 * ra + 0.02 if g[k].numel() >= 64 else 2 * ra + 0.3
            worst.append((ro / bound, k, ro, ra))
            assert ro <= bound, (k, ro, ra)
        print("batch %d: tightest gradient bound used to %.2f by %s (rel %.4f, autocast %.4f)" % ((batch,) + max(worst)))
        assert _rel(g["conv_p2.bias"], grads_t["conv_p2.bias"]) <= 1e-2            # softmax - pi summed over squares: no deep bf16 chain
        stats = tr.get_tensors(_lib.TRAIN_PARAMS, list(bn_t))
        for k in bn_t:
            assert _rel(stats[k], bn_t[k]) <= 5e-3, k
    finally:
        tr.close()


def test_adam_update_is_torch_adam(setup):
    """torch.optim.Adam fed with the trainer's own gradients must produce the trainer's weights: isolates the optimiser arithmetic"""
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.trainer import Trainer
    eng, dev = setup
    rec = _records(32, seed=4)
    rows = np.arange(16, dtype=np.int32)
    model = _model(1).to(dev)
    ref_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr = Trainer(eng, model, batch_size=16)
    tr.set_records(rec)
    tr.step(rows, flags=_lib.TRAIN_NO_UPDATE)
    g = tr.get_tensors(_lib.TRAIN_GRADS)
    tr.close()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    for k, p in model.named_parameters():
        p.grad = g[k].view(p.shape).to(dev)
    opt.step()
    want = {k: p.detach().cpu().clone() for k, p in model.named_parameters()}
    model.load_state_dict(ref_sd)
    tr = Trainer(eng, model, batch_size=16)
    tr.set_records(rec)
    tr.step(rows)
    got = tr.get_tensors(_lib.TRAIN_PARAMS, tr.param_keys)
    m = tr.get_tensors(_lib.TRAIN_EXP_AVG)
    assert tr.step_count == 1
    tr.close()
    for k in want:
        assert (got[k] - want[k].flatten()).abs().max().item() <= 1e-7, k
        assert (m[k] - 0.1 * (g[k] + 1e-4 * ref_sd[k].flatten().cpu())).abs().max().item() <= 1e-6 * max(1.0, g[k].abs().max().item()), k


def test_loss_trajectory_tracks_torch_and_step_lr(setup):
    """25 optimiser steps from the same start: the cross entropy follows torch's fp32 run within 0.05 nat at every step and falls;
    StepLR(5, 0.5) is honoured (the update after the decay is half as large)"""
    from sigma_zero_b200.trainer import Trainer
    eng, dev = setup
    rec = _records(128, seed=5)
    rng = np.random.default_rng(6)
    batches = [rng.permutation(128)[:32].astype(np.int32) for _ in range(25)]
    model = _model(2).to(dev).train()
    ref_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=500, gamma=0.95)
    traj_t = []
    for rows in batches:
        x, pi, z = _torch_batch(rec, rows, dev)
        m_, c_, _, _ = _torch_step(model, x, pi, z)
        opt.step(); sched.step()
        traj_t.append((m_, c_))
    model.load_state_dict(ref_sd)
    tr = Trainer(eng, model, batch_size=32)
    tr.set_records(rec)
    traj_o = [tr.step(rows) for rows in batches]
    tr.close()
    for (mt, ct), (mo, co) in zip(traj_t, traj_o):
        assert abs(co - ct) <= 0.05 and abs(mo - mt) <= 0.05, (traj_t, traj_o)
    assert traj_o[-1][1] < traj_o[0][1] - 1.0
    # StepLR: with lr_step = 1, gamma = 0.5 the second update is taken at half the first one's learning rate; the first two Adam
    # updates are sign-like (|m / sqrt(v)| ~ 1), so the parameter movement halves
    tr = Trainer(eng, model, batch_size=32, lr_step=1, lr_gamma=0.5, weight_decay=0.0)
    tr.set_records(rec)
    name = "resnet_blocks.18.conv2.weight"
    w0 = tr.get_tensors(0, [name])[name]
    tr.step(batches[0])
    w1 = tr.get_tensors(0, [name])[name]
    tr.step(batches[0])
    w2 = tr.get_tensors(0, [name])[name]
    tr.close()
    d1, d2 = (w1 - w0).abs().median().item(), (w2 - w1).abs().median().item()
    assert abs(d1 - 1e-4) <= 2e-6 and 0.3 <= d2 / d1 <= 0.7, (d1, d2)


def test_deterministic_resumable_and_graph_free_path_identical(setup):
    """(a) two trainers given the same steps end with bit-identical weights (fixed-order reductions, no atomics on values);
    (b) 6 steps == 3 steps, state exported through torch's optimiser format into a NEW trainer, 3 more steps -- bit for bit;
    (c) the captured CUDA graph and plain launches (SZB_TRAIN_NO_GRAPH=1) are the same computation"""
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.trainer import Trainer
    eng, dev = setup
    rec = _records(96, seed=7)
    rng = np.random.default_rng(8)
    batches = [rng.permutation(96)[:24].astype(np.int32) for _ in range(6)]
    model = _model(3)

    def run(split=None, no_graph=False):
        m = _model(3)
        if no_graph:
            os.environ["SZB_TRAIN_NO_GRAPH"] = "1"
        try:
            tr = Trainer(eng, m, batch_size=24)
        finally:
            os.environ.pop("SZB_TRAIN_NO_GRAPH", None)
        tr.set_records(rec)
        for i, rows in enumerate(batches):
            if split is not None and i == split:
                tr.write_back(m, steps_taken=i)
                state = tr.optimiser_state(m)
                tr.close()
                tr = Trainer(eng, m, batch_size=24, step0=i)
                tr.load_optimiser_state(state)
                tr.set_records(rec)
            tr.step(rows, want_losses=(i % 2 == 0))
        out = tr.get_tensors(_lib.TRAIN_PARAMS)
        tr.close()
        return out

    a, b, c, d = run(), run(), run(split=3), run(no_graph=True)
    for k in a:
        assert torch.equal(a[k], b[k]), "not deterministic: " + k
        assert torch.equal(a[k], c[k]), "resume differs: " + k
        assert torch.equal(a[k], d[k]), "graph and plain launches differ: " + k
    assert not torch.equal(a["conv1.weight"], model.state_dict()["conv1.weight"].flatten())


def test_reference_batch_size(setup):
    """batch 128 (train_RL.py:175) and an odd 255: the BatchNorm kernels run their widest grids (every block must be resident for the grid
    barrier), steps are queued without reading losses back; the losses fall and two runs agree bit for bit"""
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.trainer import Trainer
    eng, dev = setup
    rec = _records(512, seed=11)
    rng = np.random.default_rng(12)
    for batch in (128, 255):
        batches = [rng.permutation(512)[:batch].astype(np.int32) for _ in range(12)]
        outs = []
        for _ in range(2):
            tr = Trainer(eng, _model(6), batch_size=batch)
            tr.set_records(rec)
            first = tr.step(batches[0])
            for rows in batches[1:-1]:
                tr.step(rows, want_losses=False)
            last = tr.step(batches[-1])
            hist = tr.loss_history()                        # the device-side ring: what a loop that never synchronises reads at the end
            assert len(hist) == 12 and hist[0] == first and hist[-1] == last and np.isfinite(hist).all()
            assert tr.loss_history(3, 2) == hist[3:5]
            outs.append((first, last, tr.get_tensors(_lib.TRAIN_PARAMS, ["conv1.weight", "fc_v2.weight", "resnet_blocks.9.bn1.running_var"])))
            tr.close()
        (f0, l0, w0), (f1, l1, w1) = outs
        assert f0 == f1 and l0 == l1 and all(torch.equal(w0[k], w1[k]) for k in w0)
        assert np.isfinite([*f0, *l0]).all() and l0[1] < f0[1] - 0.3, (f0, l0)


def test_train_on_records_product_path(setup):
    """train_RL.train_on_records on a CUDA device runs the library's trainer: weights change in the torch module, the optimiser and
    scheduler objects carry the state on (interchangeable with the torch trainer's checkpoints), the inference network of the same
    context already holds the trained weights, and a second call continues where the first stopped"""
    from sigma_zero_b200 import runtime
    from sigma_zero_b200.train_RL import make_optimiser, train_on_records
    rec = _records(64, seed=9)
    model = _model(4)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    opt, sched = make_optimiser(model)
    hist = train_on_records(model, rec, epochs=2, batch_size=16, optimiser=opt, lr_scheduler=sched, device="cuda", seed=1)
    assert len(hist) == 8 and all(np.isfinite(h).all() for h in hist)
    sd = {k: v.cpu() for k, v in model.state_dict().items()}       # the module now lives on the training device, as with the torch path
    assert not torch.equal(sd["conv1.weight"], before["conv1.weight"]) and not torch.equal(sd["norm_layer.running_mean"], before["norm_layer.running_mean"])
    assert int(sd["norm_layer.num_batches_tracked"]) == int(before["norm_layer.num_batches_tracked"]) + 8
    st = opt.state_dict()
    assert len(st["state"]) == 129 and float(st["state"][0]["step"]) == 8 and sched.last_epoch == 8
    assert not model.training
    eng = runtime.get_engine()
    digest_trained = eng.net_checksum()
    eng.load_state_dict(model.state_dict())                   # the same weights through the ordinary host path
    assert eng.net_checksum() == digest_trained
    # a second call resumes: equal to one 16-step run made by a fresh trainer fed the same batches
    hist2 = train_on_records(model, rec, epochs=2, batch_size=16, optimiser=opt, lr_scheduler=sched, device="cuda", seed=2)
    assert float(opt.state_dict()["state"][0]["step"]) == 16 and sched.last_epoch == 16
    model_b = _model(4)
    opt_b, sched_b = make_optimiser(model_b)
    train_on_records(model_b, rec, epochs=2, batch_size=16, optimiser=opt_b, lr_scheduler=sched_b, device="cuda", seed=1)
    hist2_b = train_on_records(model_b, rec, epochs=2, batch_size=16, optimiser=opt_b, lr_scheduler=sched_b, device="cuda", seed=2)
    assert hist2 == hist2_b
    for k, v in model.state_dict().items():
        assert torch.equal(v, model_b.state_dict()[k]), k


def test_bad_calls_are_refused(setup):
    from sigma_zero_b200._lib import SzbError
    from sigma_zero_b200.trainer import Trainer
    eng, dev = setup
    rec = _records(16, seed=10)
    tr = Trainer(eng, _model(5), batch_size=8)
    try:
        with pytest.raises(SzbError):
            tr.step(np.arange(4, dtype=np.int32))                           # no records yet
        tr.set_records(rec)
        with pytest.raises(SzbError):
            tr.step(np.arange(1, dtype=np.int32))                           # BatchNorm needs two boards
        with pytest.raises(SzbError):
            tr.step(np.arange(9, dtype=np.int32))                           # more than the trainer was created for
        with pytest.raises(SzbError):
            tr.step(np.array([0, 1, 2, 16], dtype=np.int32))                # row outside the records
        with pytest.raises(SzbError):
            tr.get_tensors(0, ["no.such.tensor"], shapes={"no.such.tensor": (1,)})
        mse, ce = tr.step(np.arange(8, dtype=np.int32))                     # and the trainer still works afterwards
        assert np.isfinite([mse, ce]).all()
        # replacing the records invalidates the captured steps (they carry the old buffers' addresses): the same rows of NEW records
        # must give what a fresh trainer in the same state gives
        from sigma_zero_b200 import _lib
        rec2 = _records(24, seed=13)
        tr.set_records(rec2)
        got = tr.step(np.arange(8, dtype=np.int32), flags=_lib.TRAIN_FORWARD_ONLY)
        m = _model(5)
        tr.write_back(m)
    finally:
        tr.close()
    tr = Trainer(eng, m, batch_size=8)
    try:
        tr.set_records(rec2)
        assert tr.step(np.arange(8, dtype=np.int32), flags=_lib.TRAIN_FORWARD_ONLY) == got
    finally:
        tr.close()
    # a trainer that was given only part of the state_dict refuses to step and names a missing tensor
    import ctypes
    cfg = _lib.TrainConfig(8, 1e-4, 0.9, 0.999, 1e-8, 1e-4, 500, 0.95, 0.1, 1e-5, 0, 0, 0)
    eng._check(eng.lib.szb_train_create(eng._h, ctypes.byref(cfg)))
    rows = np.arange(4, dtype=np.int32)
    rc = eng.lib.szb_train_step(eng._h, 4, rows.ctypes.data_as(ctypes.c_void_p), 0, None)
    assert rc == -5 and b"never set" in eng.lib.szb_last_error(eng._h)
    eng.lib.szb_train_destroy(eng._h)
