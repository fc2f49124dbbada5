"""Runs a few steps of the CUDA trainer on synthetic records (measurement aid).  Under ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
        python scripts/trainer_profile.py --steps 1
captures exactly the launches of the steps after the warm-up (cudaProfilerStart / Stop bracket them); scripts/trainer_profile.py
--summarise <csv> folds such a launch list per kernel."""
import argparse, collections, csv, os, sys, time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def summarise(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    head = rows[0]
    name_i, val_i = head.index("Kernel Name"), head.index("Metric Value")
    unit_i = head.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        ns = float(r[val_i].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[unit_i], 1)
        k = r[name_i].split("(")[0]
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1; t[1] += ns
    total = sum(v[1] for v in tot.values())
    print("%-40s %6s %10s %8s %6s" % ("kernel", "count", "total us", "avg us", "share"))
    for k, (c, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-40s %6d %10.1f %8.2f %5.1f%%" % (k[:40], c, ns / 1e3, ns / c / 1e3, 100 * ns / total))
    print("%-40s %6d %10.1f" % ("all", sum(v[0] for v in tot.values()), total / 1e3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--summarise")
    a = ap.parse_args()
    if a.summarise:
        return summarise(a.summarise)
    import torch
    from scripts.trainer_check import make_model, synthetic_records
    from sigma_zero_b200.engine import Engine
    from sigma_zero_b200.trainer import Trainer
    B = a.batch
    rec = synthetic_records(4 * B, seed=3)
    eng = Engine(max_games=2, max_searches=8, device=0)
    tr = Trainer(eng, make_model(0), batch_size=B)
    tr.set_records(rec)
    rng = np.random.default_rng(1)
    batches = [rng.permutation(4 * B)[:B].astype(np.int32) for _ in range(8)]
    for i in range(a.warmup):
        tr.step(batches[i % 8])
    eng.synchronize()
    torch.cuda.profiler.start()
    t0 = time.perf_counter()
    for i in range(a.steps):
        tr.step(batches[i % 8], want_losses=False)
    eng.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    torch.cuda.profiler.stop()
    print("batch %d: %.3f ms per step (%.0f positions/s)" % (B, dt * 1e3, B / dt))


if __name__ == "__main__":
    main()
