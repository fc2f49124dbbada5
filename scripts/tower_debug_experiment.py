import os, sys
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
G = 1024
for dbg in ("0", "1", "2", "3"):
    os.environ["SZB_TOWER_DEBUG"] = dbg
    eng = Engine(max_games=G, max_searches=8, cohorts=1)
    eng.load_state_dict(model.state_dict())
    eng.reset([-1] * G)
    print("debug=%s  tower burst %.3f ms   one layer %.4f ms" % (dbg, eng.time_kernel(5, G, 10), eng.time_kernel(4, G, 20)))
    eng.close()
