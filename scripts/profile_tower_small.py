"""ncu target: a few whole-tower launches at a small batch (N-split path) and at the 512-board chunk of large evaluations."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
eng = Engine(max_games=512, max_searches=8)
eng.load_state_dict(model.state_dict())
eng.reset([-1] * 512)
for n in (64, 512):
    print(n, eng.time_kernel(5, n, 3))
eng.close()
