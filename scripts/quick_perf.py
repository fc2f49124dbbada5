"""Ad-hoc first performance look (not the bench contract): kernel timings on one B200."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16, EVAL_NET_FP32, EVAL_HASH
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
eng = Engine(max_games=G, max_searches=max(S, 100))
t = time.time(); eng.load_state_dict(model.state_dict()); print("load", time.time() - t)
eng.reset([-1] * G)
flop_layer = 64 * G * 256 * 2304 * 2
flop_fwd = 2914845184 * G
ms = eng.time_kernel(0, G, 20); print("tower conv bf16: %.3f ms  %.1f TFLOP/s" % (ms, flop_layer / ms / 1e9))
ms = eng.time_kernel(1, G, 5); print("forward bf16: %.3f ms  %.1f TFLOP/s  %.0f evals/s" % (ms, flop_fwd / ms / 1e9, G / ms * 1e3))
ms = eng.time_kernel(3, min(G, 256), 3); print("tower conv fp32 (n=%d): %.3f ms  %.1f TFLOP/s" % (min(G, 256), ms, 64 * min(G, 256) * 256 * 2304 * 2 / ms / 1e9))
for ev, name in ((EVAL_HASH, "hash"), (EVAL_NET_BF16, "bf16")):
    eng.reset([-1] * G)
    eng.set_profiling(True)
    eng.search(S, 2.0, True, ev, want_visits=False, want_children=False)
    t = time.time(); eng.search(S, 2.0, True, ev, want_visits=False, want_children=False); dt = time.time() - t
    pt = eng.phase_times()
    print(name, "search %d x %d sims: %.3f s  %.0f sims/s" % (G, S, dt, G * S / dt), json.dumps(pt))
    eng.set_profiling(False)
print(eng.stats())
