"""Per-phase device time of one simulation step (library profiling: CUDA events around every phase) for a few batch sizes."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
S = 200
for G in [int(x) for x in sys.argv[1:]] or [1, 8, 63, 250, 1024]:
    eng = Engine(max_games=G, max_searches=S)
    eng.load_state_dict(model.state_dict())
    eng.reset([-1] * G)
    eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    eng.set_profiling(True)
    eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    pt = eng.phase_times()
    eng.set_profiling(False)
    n = max(1, pt["steps"])
    tower = pt["conv_ms"] / max(1, pt["conv_launches"])
    print("G=%4d  select %.1f us  expand %.1f us  eval %.1f us (tower launch %.1f us)  finish %.1f us  | sum %.1f us" %
          (G, pt["select_ms"] / n * 1e3, pt["expand_ms"] / n * 1e3, pt["eval_ms"] / n * 1e3, tower * 1e3, pt["finish_ms"] / n * 1e3,
           (pt["select_ms"] + pt["expand_ms"] + pt["eval_ms"] + pt["finish_ms"]) / n * 1e3), flush=True)
    eng.close()
