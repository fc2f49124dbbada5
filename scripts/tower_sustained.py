"""Tower kernel: burst (back-to-back launches) and sustained (inside a search loop) time per launch."""
import os, sys, time
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
G, S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 96
flop = 2 * 64 * (256 * (119 * 9 + 38 * 2304 + 256) + 73 * 256) * G
for cohorts in (1, 2):
    eng = Engine(max_games=G, max_searches=S, cohorts=cohorts)
    eng.load_state_dict(model.state_dict())
    eng.reset([-1] * G)
    burst = eng.time_kernel(5, G, 10)
    eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    t = time.time()
    for _ in range(4):
        eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    dt = (time.time() - t) / 4
    eng.set_profiling(True)
    for _ in range(2):
        eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    pt = eng.phase_times()
    ms = pt["conv_ms"] / pt["conv_launches"]
    print("G=%d cohorts=%d: tower burst %.3f ms (%.0f TF)  sustained %.3f ms (%.0f TF)  search %.3f ms/step = %.0f sims/s" %
          (G, cohorts, burst, flop / burst / 1e9, ms, flop / ms / 1e9, dt / S * 1e3, G * S / dt))
    eng.close()
