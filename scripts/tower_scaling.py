"""Tower kernel time vs batch, and search throughput with 1 / 2 cohorts (measurement aid, not the bench contract)."""
import sys, time, json, os
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
flop_tower = 2 * 64 * 256 * (119 * 9 + 38 * 2304 + 256)
eng = Engine(max_games=4096, max_searches=64)
eng.load_state_dict(model.state_dict())
eng.reset([-1] * 4096)
for n in (128, 256, 296, 512, 592, 1024, 1184, 2048, 4096):
    ms = eng.time_kernel(5, n, 20)
    print("tower n=%5d  %8.3f ms  %7.1f TFLOP/s  %.3f us/board" % (n, ms, flop_tower * n / ms / 1e9, ms * 1e3 / n))
eng.close()
S = 64
for G in (1024, 2048, 4096):
    for c in (1, 2):
        eng = Engine(max_games=G, max_searches=S, cohorts=c)
        eng.load_state_dict(model.state_dict())
        eng.reset([-1] * G)
        eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
        t = time.time()
        for _ in range(3):
            eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
        dt = (time.time() - t) / 3
        print("search G=%d cohorts=%d: %.3f ms/step  %.0f sims/s" % (G, c, dt / S * 1e3, G * S / dt))
        eng.close()
