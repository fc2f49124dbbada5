"""Small-batch latency (measurement aid): tower launch time and search step time for few games, with the N split of
small batches forced off (SZB_TOWER_NSPLIT=1) and automatic.  Run each setting in its own process:
    SZB_TOWER_NSPLIT=1 python scripts/small_batch.py ; python scripts/small_batch.py"""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
tag = "nsplit=" + os.environ.get("SZB_TOWER_NSPLIT", "auto")
res = {"setting": tag, "tower_ms": {}, "search": {}}
eng = Engine(max_games=512, max_searches=64)
eng.load_state_dict(model.state_dict())
eng.reset([-1] * 512)
for n in (1, 4, 8, 16, 32, 64, 72, 100, 148, 200, 296, 512):
    ms = eng.time_kernel(5, n, 50)
    res["tower_ms"][n] = ms
    print("%s tower n=%4d  %8.4f ms  (%.2f us/layer)" % (tag, n, ms, ms * 1e3 / 41), flush=True)
res["cluster_tower_ms"] = {}
for n in (1, 2, 4, 8, 12, 15):
    try:
        ms = eng.time_kernel(6, n, 50)
    except Exception as e:                     # more boards than clusters fit at once on this device
        print("%s cluster tower n=%d: %s" % (tag, n, e), flush=True)
        continue
    res["cluster_tower_ms"][n] = ms
    print("%s cluster tower (k_tower_cl, whole forward) n=%4d  %8.4f ms  (%.2f us/layer)" % (tag, n, ms, ms * 1e3 / 41), flush=True)
eng.close()
S = 200
for G in (1, 8, 18, 63, 125, 250):
    eng = Engine(max_games=G, max_searches=S)
    eng.load_state_dict(model.state_dict())
    eng.reset([-1] * G)
    eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    t = time.time()
    for _ in range(3):
        eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    dt = (time.time() - t) / 3
    res["search"][G] = {"ms_per_step": dt / S * 1e3, "sims_per_s": G * S / dt}
    print("%s search G=%d: %.3f ms/step  %.0f sims/s" % (tag, G, dt / S * 1e3, G * S / dt), flush=True)
    eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/small_batch_%s.json" % tag.replace("=", "_"), "w"), indent=1)
