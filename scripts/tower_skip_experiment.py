"""EXPERIMENT: how much would the tower gain if activation tiles were loaded once per K-chunk instead of once per tap
(bit 0) and/or weight tiles were shared by two pairs (bit 1)?  The skip modes compute garbage; only time matters."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
G, S = 1024, 96
for skip in ("0", "1", "2", "3"):
    os.environ["SZB_TOWER_DEBUG_SKIP"] = skip
    eng = Engine(max_games=G, max_searches=S, cohorts=1)
    eng.load_state_dict(model.state_dict())
    eng.reset([-1] * G)
    burst = eng.time_kernel(5, G, 10)
    eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    eng.set_profiling(True)
    for _ in range(4):
        eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    pt = eng.phase_times()
    print("skip=%s  tower burst %.3f ms   sustained %.3f ms per launch (%d launches)" % (skip, burst, pt["conv_ms"] / pt["conv_launches"], pt["conv_launches"]))
    eng.close()
