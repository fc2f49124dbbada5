"""GPU diagnostics of the CUDA trainer (szb_train_*) against torch autograd on the same GPU: forward (losses, logits, value), every
gradient tensor, one Adam update, a short loss trajectory, and step time next to torch eager.  Prints one line per check and writes
gpurun_out/trainer_check.json.  torch's own bf16-autocast run is the yardstick for what bf16 operands cost against fp32.

    python scripts/trainer_check.py [--batch 128] [--steps 20] [--probe]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma_zero_b200 import _lib, records as records_mod
from sigma_zero_b200.engine import Engine
from sigma_zero_b200.network import policyNN
from sigma_zero_b200.trainer import Trainer


def synthetic_records(n, seed=0):
    rng = np.random.default_rng(seed)
    bits = rng.random((n, 119, 64)) < 0.12
    states = np.packbits(bits, axis=-1, bitorder="little").view("<u8").reshape(n, 119)
    idx, prob, off = [], [], [0]
    for i in range(n):
        k = int(rng.integers(5, 45))
        ind = np.sort(rng.choice(4672, size=k, replace=False))
        p = rng.random(k).astype(np.float32) ** 3
        p /= p.sum()
        idx.extend(ind.tolist())
        prob.extend(p.tolist())
        off.append(len(idx))
    return {"states": states.astype(np.uint64), "pi_index": np.array(idx, np.uint16), "pi_prob": np.array(prob, np.float32),
            "pi_off": np.array(off, np.int64), "z": rng.integers(-1, 2, n).astype(np.int8), "colour": np.ones(n, bool), "game": np.zeros(n, np.int32)}


def make_model(seed=0):
    torch.manual_seed(seed)
    model = policyNN({})
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.copy_(0.5 + torch.rand(mod.weight.shape, generator=g))
                mod.bias.copy_(0.2 * torch.randn(mod.bias.shape, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(mod.running_mean.shape, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.running_var.shape, generator=g))
    return model


def torch_batch(rec, rows, dev):
    x = torch.from_numpy(records_mod.unpack_states(rec, rows)).to(device=dev, dtype=torch.float32)
    pi = torch.from_numpy(records_mod.dense_policy(rec, rows)).to(dev)
    z = torch.from_numpy(rec["z"][rows].astype(np.float32)).to(dev)
    return x, pi, z


def torch_step(model, x, pi, z, autocast=False):
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        p, v = model.forward_torch(x)
    p, v = p.float(), v.float()
    mse = torch.nn.functional.mse_loss(v.squeeze(-1), z)
    ce = torch.nn.functional.cross_entropy(p, pi)
    (mse + ce).backward()
    return float(mse.detach()), float(ce.detach()), p.detach(), v.detach().squeeze(-1)


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    nb = float(b.norm())
    return float((a - b).norm() / (nb if nb > 0 else 1.0)), float((a @ b) / ((float(a.norm()) * nb) or 1.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--probe", action="store_true", help="try alternative wgrad descriptor strides and report each")
    ap.add_argument("--time-steps", type=int, default=30)
    ap.add_argument("--out", default="gpurun_out/trainer_check.json")
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    B = a.batch
    rec = synthetic_records(4 * B, seed=3)
    rows = np.arange(B, dtype=np.int32) * 3 % (4 * B)
    report = {"batch": B}

    model = make_model(0).to(dev).train()
    ref_sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, pi, z = torch_batch(rec, rows, dev)
    mse_t, ce_t, p_t, v_t = torch_step(model, x, pi, z)
    grads_t = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    bn_after = {k: v.detach().clone() for k, v in model.state_dict().items() if "running" in k}
    model.load_state_dict(ref_sd)
    mse_a, ce_a, p_a, v_a = torch_step(model, x, pi, z, autocast=True)
    grads_a = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    model.load_state_dict(ref_sd)
    print("torch fp32: mse %.6f ce %.6f | torch bf16 autocast: mse %.6f ce %.6f" % (mse_t, ce_t, mse_a, ce_a))

    eng = Engine(max_games=2, max_searches=8, device=0)
    probes = [(0, 0)]
    if a.probe:
        probes += [(1024, 8192), (8192, 128), (128, 8192)]          # all wrong, on purpose: (8192, 1024) is the layout TMA writes
    for probe in probes:
        tr = Trainer(eng, model, batch_size=B, _probe=probe)
        tr.set_records(rec)
        t0 = time.time()
        mse_o, ce_o = tr.step(rows, flags=_lib.TRAIN_NO_UPDATE)
        print("probe lbo/sbo %s: ours mse %.6f ce %.6f (first step %.1f ms)" % (probe, mse_o, ce_o, (time.time() - t0) * 1e3))
        lg, val = tr.activations(B)
        e_lg, e_lg_a = rel(lg, p_t), rel(p_a, p_t)
        e_v, e_v_a = rel(val, v_t), rel(v_a, v_t)
        print("  logits rel %.4f cos %.5f (torch autocast rel %.4f) | value rel %.4f (autocast %.4f)" % (e_lg[0], e_lg[1], e_lg_a[0], e_v[0], e_v_a[0]))
        shapes = {k: tuple(p.shape) for k, p in model.named_parameters()}
        g = tr.get_tensors(_lib.TRAIN_GRADS)
        rows_out = []
        for k in tr.param_keys:
            ro, co = rel(g[k], grads_t[k])
            ra, ca = rel(grads_a[k], grads_t[k])
            rows_out.append((k, ro, co, ra, ca))
        worst = sorted(rows_out, key=lambda r: -r[1])
        for k, ro, co, ra, ca in (rows_out if probe == (0, 0) else worst[:6]):
            print("  grad %-38s rel %.4f cos %.5f | autocast rel %.4f cos %.5f%s" % (k, ro, co, ra, ca, "   <<<" if ro > max(0.08, 3 * ra) else ""))
        if probe == (0, 0):
            report.update(mse_torch=mse_t, ce_torch=ce_t, mse_ours=mse_o, ce_ours=ce_o, logits_rel=e_lg[0], logits_rel_autocast=e_lg_a[0],
                          value_rel=e_v[0], grads={k: {"rel": ro, "cos": co, "autocast_rel": ra} for k, ro, co, ra, ca in rows_out})
            sd_o = tr.get_tensors(_lib.TRAIN_PARAMS, [k for k in tr.keys if "running" in k])
            bad = [(k, rel(sd_o[k], bn_after[k])[0]) for k in sd_o]
            print("  running statistics: worst rel %.5f (%s)" % (max(b[1] for b in bad), max(bad, key=lambda b: b[1])[0]))
            report["running_stats_worst_rel"] = max(b[1] for b in bad)
        tr.close()

    # ---- one Adam update against torch.optim.Adam fed with OUR gradients (isolates the optimiser arithmetic) -----------------
    tr = Trainer(eng, model, batch_size=B)
    tr.set_records(rec)
    tr.step(rows, flags=_lib.TRAIN_NO_UPDATE)
    g = tr.get_tensors(_lib.TRAIN_GRADS)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    for k, p in model.named_parameters():
        p.grad = g[k].view(p.shape).to(dev)
    opt.step()
    want = {k: p.detach().clone() for k, p in model.named_parameters()}
    model.load_state_dict(ref_sd)
    tr.close()
    tr = Trainer(eng, model, batch_size=B)
    tr.set_records(rec)
    tr.step(rows)
    got = tr.get_tensors(_lib.TRAIN_PARAMS, tr.param_keys)
    worst = max(float((got[k].view(-1) - want[k].cpu().view(-1)).abs().max()) for k in tr.param_keys)
    print("Adam: worst |w_ours - w_torch| after one step %.3e (lr 1e-4)" % worst)
    report["adam_worst_abs"] = worst
    tr.close()

    # ---- loss trajectory ----------------------------------------------------------------------------------------------------
    rng = np.random.default_rng(5)
    batches = [rng.permutation(4 * B)[:B].astype(np.int32) for _ in range(a.steps)]
    model.load_state_dict(ref_sd)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=500, gamma=0.95)
    traj_t = []
    for rws in batches:
        xb, pb, zb = torch_batch(rec, rws, dev)
        m_, c_, _, _ = torch_step(model, xb, pb, zb)
        opt.step(); sched.step()
        traj_t.append((m_, c_))
    model.load_state_dict(ref_sd)
    tr = Trainer(eng, model, batch_size=B)
    tr.set_records(rec)
    traj_o = [tr.step(rws) for rws in batches]
    for i in range(0, a.steps, max(1, a.steps // 10)):
        print("  step %3d: torch mse %.5f ce %.5f | ours mse %.5f ce %.5f" % (i, traj_t[i][0], traj_t[i][1], traj_o[i][0], traj_o[i][1]))
    report["trajectory_torch"], report["trajectory_ours"] = traj_t, traj_o
    # trained weights go back into a torch module and must load into the inference network
    m2 = make_model(0)
    tr.write_back(m2, steps_taken=a.steps)
    keys, numels, flat = tr.flat_weights()
    eng.load_flat_device(keys, numels, flat)
    print("write_back + load_flat_device ok; digest %x" % eng.net_checksum())

    # ---- timing ---------------------------------------------------------------------------------------------------------------
    torch.cuda.synchronize()
    for _ in range(3):
        tr.step(batches[0], want_losses=False)
    eng.synchronize()
    t0 = time.perf_counter()
    for i in range(a.time_steps):
        tr.step(batches[i % len(batches)], want_losses=False)
    eng.synchronize()
    ours_ms = (time.perf_counter() - t0) / a.time_steps * 1e3
    tr.close()

    def torch_time(autocast, tf32):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        model.load_state_dict(ref_sd)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)
        xb, pb, zb = torch_batch(rec, batches[0], dev)
        for _ in range(3):
            torch_step(model, xb, pb, zb, autocast); opt.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.time_steps):
            model.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                p, v = model.forward_torch(xb)
            loss = torch.nn.functional.mse_loss(v.float().squeeze(-1), zb) + torch.nn.functional.cross_entropy(p.float(), pb)
            loss.backward(); opt.step()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / a.time_steps * 1e3

    t_fp32, t_tf32, t_bf16 = torch_time(False, False), torch_time(False, True), torch_time(True, True)
    print("step time, batch %d: ours %.3f ms | torch eager fp32 %.3f ms, tf32 (torch default for convolutions) %.3f ms, bf16 autocast %.3f ms (inputs resident)"
          % (B, ours_ms, t_fp32, t_tf32, t_bf16))
    report.update(step_ms_ours=ours_ms, step_ms_torch_fp32=t_fp32, step_ms_torch_tf32=t_tf32, step_ms_torch_bf16_autocast=t_bf16)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
