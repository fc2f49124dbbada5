"""A/B of the small-batch towers inside a search (same box, alternating processes are not needed: the choice is per context):
SZB_TOWER_CLUSTER=0 (CTA-pair kernel with N split) vs default (cluster-resident kernel up to 15 boards)."""
import os, sys, time, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ".")
    from oracle import ref_path
    from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
    torch.manual_seed(0)
    sd = ref_path.build_policy_nn().eval().state_dict()
    out = {}
    for G in (1, 2, 4, 8, 15):
        eng = Engine(max_games=G, max_searches=400)
        eng.load_state_dict(sd)
        eng.reset([-1] * G)
        eng.search(400, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
        t = time.time()
        for _ in range(3):
            eng.search(400, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
        out[G] = (time.time() - t) / 3 / 400 * 1e3
        eng.close()
    print(json.dumps(out))
else:
    for rep in range(2):
        for tag, env in (("pair kernel", {"SZB_TOWER_CLUSTER": "0"}), ("cluster kernel", {})):
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
            print("%-15s ms/step %s" % (tag, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]), flush=True)
