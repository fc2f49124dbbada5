"""Ad-hoc kernel timings on one B200 (not the bench contract): tower kernel variants, whole forward, search phases."""
import sys, time, json, os
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16, EVAL_HASH
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 50
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
flop_layer = 64 * G * 256 * 2304 * 2
flop_tower = 2 * 64 * 256 * (119 * 9 + 38 * 2304 + 256) * G
flop_fwd = 2914845184 * G
for mode in sys.argv[3].split(",") if len(sys.argv) > 3 else ("0", "1", "2"):
    os.environ["SZB_TOWER_MODE"] = mode
    eng = Engine(max_games=G, max_searches=max(S, 100))
    eng.load_state_dict(model.state_dict())
    eng.reset([-1] * G)
    print("== SZB_TOWER_MODE", mode)
    for which, name, fl in ((0, "1-CTA layer", flop_layer), (4, "pair layer", flop_layer), (5, "pair tower (1 launch)", flop_tower), (1, "whole forward", flop_fwd)):
        ms = eng.time_kernel(which, G, 10)
        print("  %-22s %8.3f ms  %7.1f TFLOP/s" % (name, ms, fl / ms / 1e9))
    eng.set_profiling(True)
    eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    t = time.time(); eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False); dt = time.time() - t
    print("  search %d x %d sims: %.3f s  %.0f sims/s" % (G, S, dt, G * S / dt), json.dumps(eng.phase_times()))
    eng.close()
