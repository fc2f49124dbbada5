"""Per-layer timeline of the cluster-resident tower (k_tower_cl), cluster 0 / CTA 0 (measurement aid).
Usage: python scripts/cluster_trace.py [boards=1]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
eng = Engine(max_games=32, max_searches=8)
eng.load_state_dict(ref_path.build_policy_nn().eval().state_dict())
eng.reset([-1] * 32)
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/cluster_trace_%d.csv" % n
if os.path.exists(path):
    os.remove(path)
os.environ["SZB_TOWER_TRACE"] = path
ms = eng.time_kernel(6, n, 20)
eng.close()
print("".join(l for l in open(path) if l.startswith("#")))
t = np.genfromtxt(path, delimiter=",", skip_header=1, comments="#")[:, 2:]
lay = t[2:39]                                 # 3x3 tower layers
print("k_tower_cl, %d board(s): %.1f us per launch" % (n, ms * 1e3))
print("per 3x3 layer (us): input->mma issued %.2f | mma issued->acc ready %.2f | tmem read %.2f | stores %.2f | arrive %.2f | arrive->next input ready %.2f | period %.2f" % (
    (lay[:, 1] - lay[:, 0]).mean() / 1e3, (lay[:, 2] - lay[:, 1]).mean() / 1e3, (lay[:, 3] - lay[:, 2]).mean() / 1e3,
    (lay[:, 4] - lay[:, 3]).mean() / 1e3, (lay[:, 5] - lay[:, 4]).mean() / 1e3, (t[3:40, 0] - lay[:, 5]).mean() / 1e3,
    np.diff(t[2:40, 0]).mean() / 1e3))
