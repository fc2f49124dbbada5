"""Same-box A/B: search throughput at 1024 games with one cohort vs two, alternating, several repetitions."""
import sys, time
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
G, S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 200
engs = {}
for c in (1, 2):
    e = Engine(max_games=G, max_searches=S, cohorts=c)
    e.load_state_dict(model.state_dict())
    e.reset([-1] * G)
    e.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
    engs[c] = e
for rep in range(4):
    for c in (1, 2):
        t = time.time()
        engs[c].search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
        dt = time.time() - t
        print("rep %d cohorts=%d: %.3f ms/step  %.0f sims/s" % (rep, c, dt / S * 1e3, G * S / dt))
