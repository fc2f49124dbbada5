import sys
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
eng = Engine(max_games=1024, max_searches=8, cohorts=1)
eng.load_state_dict(model.state_dict())
eng.reset([-1] * 1024)
print("one layer %.4f ms, tower %.3f ms" % (eng.time_kernel(4, 1024, 3), eng.time_kernel(5, 1024, 2)))
eng.close()
