"""Where a k_tree_step goes (measurement aid): device timestamps of the first tree's warp at the phase boundaries of every step of
one search, averaged.  Usage: python scripts/step_trace.py [games ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
S = 200
os.makedirs("gpurun_out", exist_ok=True)
for G in [int(x) for x in sys.argv[1:]] or [1, 63]:
    eng = Engine(max_games=G, max_searches=S)
    eng.load_state_dict(model.state_dict())
    eng.reset([-1] * G)
    eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    path = "gpurun_out/step_trace_%d.csv" % G
    os.environ["SZB_STEP_TRACE"] = path
    eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    del os.environ["SZB_STEP_TRACE"]
    eng.close()
    rows = np.genfromtxt(path, delimiter=",", skip_header=1)[20:-1]
    names = ["tables", "finish", "select", "move_made", "movegen", "planes_mask", "input_rows"]
    t = rows[:, 1:]
    ok = (t > 0).all(axis=1)
    t = t[ok]
    d = np.diff(t, axis=1) / 1e3
    period = np.diff(rows[:, 1]) / 1e3
    print("G=%d  step period %.1f us; inside k_tree_step (us): %s | kernel total %.1f" %
          (G, period.mean(), "  ".join("%s %.1f" % (n, x) for n, x in zip(names, d.mean(axis=0))), (t[:, -1] - t[:, 0]).mean() / 1e3), flush=True)
