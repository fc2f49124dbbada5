"""Debug: per-cohort phase timestamps of one search (SZB_TRACE=<csv>), to see whether the two cohorts overlap."""
import os, sys
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
G, S = 1024, 48
eng = Engine(max_games=G, max_searches=S, cohorts=2)
eng.load_state_dict(model.state_dict())
eng.reset([-1] * G)
eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
os.environ["SZB_TRACE"] = "gpurun_out/trace.csv"
eng.set_profiling(True)
eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
eng.close()
rows = [l.strip().split(",") for l in open("gpurun_out/trace.csv")][1:]
for r in rows[40:56]:
    print(r)
