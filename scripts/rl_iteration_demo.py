"""One outer iteration of the reference's train_RL.main (train_RL.py:205-264) with its OWN arguments (:166-176: 100 searches per move,
batch 128, Chess960) on one GPU: self-play of `--games` games to their end on the engine, then `--passes` passes of fine-tuning over the
recorded positions on the library's trainer, then a short second self-play with the new weights.  Prints one JSON line with where the
time went (measurement aid; the numbers of a run are kept in profiles/).
    python scripts/rl_iteration_demo.py [--games 500] [--searches 100] [--passes 7] [--max-plies N]"""
import argparse, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma_zero_b200 import runtime
from sigma_zero_b200.network import policyNN
from sigma_zero_b200.train_RL import make_optimiser, selfplay_iteration_records, train_on_records

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=500)
ap.add_argument("--searches", type=int, default=100)
ap.add_argument("--passes", type=int, default=7)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--max-plies", type=int, default=None)
a = ap.parse_args()
torch.manual_seed(0)
model = policyNN({}).eval()
args = {"C": 2, "num_searches": a.searches, "num_selfPlay_iterations": a.games, "chess960": True, "batch_size": a.batch}
t0 = time.perf_counter()
rec, counters = selfplay_iteration_records(model, args, seed=1, max_plies=a.max_plies)
t1 = time.perf_counter()
opt, sched = make_optimiser(model)
hist = train_on_records(model, rec, epochs=a.passes, batch_size=a.batch, optimiser=opt, lr_scheduler=sched, device="cuda", seed=1)
torch.cuda.synchronize()
t2 = time.perf_counter()
rec2, _ = selfplay_iteration_records(model, dict(args, num_selfPlay_iterations=min(a.games, 64)), seed=2, max_plies=8)
t3 = time.perf_counter()
n = int(len(rec["z"]))
k = max(1, len(hist) // 20)
print(json.dumps({
    "workload": "train_RL.main, one outer iteration: %d Chess960 games x %d searches per move to the end of every game, then %d passes of "
                "fine-tuning at batch %d (loss = MSE + soft-target CE, Adam 1e-4 wd 1e-4, StepLR(500, 0.95)), 1 GPU" % (a.games, a.searches, a.passes, a.batch),
    "positions": n, "selfplay_s": t1 - t0, "selfplay_positions_per_s": n / (t1 - t0), "simulations_per_s": n * a.searches / (t1 - t0),
    "train_s": t2 - t1, "optimiser_steps": len(hist), "trained_positions_per_s": len(hist) * a.batch / (t2 - t1),
    "train_share_of_iteration": (t2 - t1) / (t2 - t0),
    "loss_mean_first_steps": [float(np.mean([h[0] for h in hist[:k]])), float(np.mean([h[1] for h in hist[:k]]))],
    "loss_mean_last_steps": [float(np.mean([h[0] for h in hist[-k:]])), float(np.mean([h[1] for h in hist[-k:]]))],
    "lr_after": opt.param_groups[0]["lr"], "adam_step": float(opt.state_dict()["state"][0]["step"]),
    "next_selfplay_with_new_weights_s": t3 - t2, "next_selfplay_positions": int(len(rec2["z"])),
    "weights": "random-init torch.manual_seed(0) (the reference's checkpoint is a git-LFS pointer)",
    "note": "train_s includes the one-off upload of the records, the trainer's set-up (allocation, weight import, graph capture) and the hand-over of "
            "weights and optimiser state back to torch",
}))
