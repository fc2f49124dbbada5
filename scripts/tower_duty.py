"""k_tower_tc2 back to back for seconds (100 % duty, power-capped) vs a short burst: what the tower kernel alone sustains."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
flop = 2 * 64 * (256 * (119 * 9 + 38 * 2304 + 256) + 73 * 256)
eng = Engine(max_games=1024, max_searches=8)
eng.load_state_dict(model.state_dict())
eng.reset([-1] * 1024)
print("-- fresh context: input planes all zero (constant activations: little switching, little power -- NOT a sustained figure)")
for n, iters in ((512, 10), (512, 3000)):
    ms = eng.time_kernel(5, n, iters)
    print("tower n=%4d x %4d launches back to back: %.4f ms per launch  %.0f TFLOP/s" % (n, iters, ms, flop * n / ms / 1e9), flush=True)
from sigma_zero_b200.engine import EVAL_NET_BF16
eng.search(8, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
print("-- after a search: the activation buffers hold real positions")
for n, iters in ((512, 10), (512, 3000), (1024, 10), (1024, 1500), (512, 3000)):
    ms = eng.time_kernel(5, n, iters)
    print("tower n=%4d x %4d launches back to back: %.4f ms per launch  %.0f TFLOP/s" % (n, iters, ms, flop * n / ms / 1e9), flush=True)
eng.close()
