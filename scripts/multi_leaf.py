"""Opt-in multi-leaf mode (virtual loss, leaves_per_tree = K): search time for few games (measurement aid)."""
import sys, time
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
S = 800
for G in (1, 8, 63):
    for K in (1, 2, 4, 8):
        eng = Engine(max_games=G, max_searches=S, leaves_per_tree=K)
        eng.load_state_dict(model.state_dict())
        eng.reset([-1] * G)
        eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
        t = time.time()
        for _ in range(2):
            eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
        dt = (time.time() - t) / 2
        print("G=%3d leaves_per_tree=%d: %.1f ms per %d-simulation move  %.0f sims/s" % (G, K, dt * 1e3, S, G * S / dt), flush=True)
        eng.close()
