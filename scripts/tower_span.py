"""Where the towers of a pipelined (two-cohort) search really run: device start / end time of every tower launch (SZB_TOWER_SPAN)."""
import os, sys, csv
os.environ["SZB_TOWER_SPAN"] = "gpurun_out/tower_span.csv"
import torch
sys.path.insert(0, ".")
from oracle import ref_path
from sigma_zero_b200.engine import Engine, EVAL_NET_BF16
torch.manual_seed(0)
model = ref_path.build_policy_nn().eval()
G, S = 1024, 800
for cohorts in ([int(x) for x in sys.argv[1:]] or (2, 1)):
    os.environ["SZB_TOWER_SPAN"] = "gpurun_out/tower_span_c%d.csv" % cohorts
    eng = Engine(max_games=G, max_searches=S, cohorts=cohorts)
    eng.load_state_dict(model.state_dict())
    eng.reset([-1] * G)
    import time
    t_last = 0
    for _ in range(3):                                   # 3 x 800 steps x 2 launches: the last ~3000 launches are in steady state
        t0 = time.time()
        eng.search(S, 2.0, True, EVAL_NET_BF16, want_visits=False, want_children=False)
        t_last = time.time() - t0
    print("cohorts=%d pairs=%s exclusive=%s: %.0f sims/s (last search)" % (cohorts, os.environ.get("SZB_TOWER_PAIRS", "74"), os.environ.get("SZB_TOWER_EXCLUSIVE", "0"), G * S / t_last))
    eng.close()
    rows = list(csv.DictReader(open("gpurun_out/tower_span_c%d.csv" % cohorts)))[-2000:]
    dur = [(int(r["end_ns"]) - int(r["start_ns"])) / 1e3 for r in rows]
    gap = [(int(b["start_ns"]) - int(a["end_ns"])) / 1e3 for a, b in zip(rows, rows[1:])]
    period = (int(rows[-1]["start_ns"]) - int(rows[0]["start_ns"])) / 1e3 / (len(rows) - 1)
    dur.sort(); gap.sort()
    print("cohorts=%d: tower launch duration median %.1f us (p10 %.1f, p90 %.1f); gap to the next launch median %.1f us (p10 %.1f, p90 %.1f); "
          "launch period %.1f us" % (cohorts, dur[len(dur) // 2], dur[len(dur) // 10], dur[9 * len(dur) // 10],
                                     gap[len(gap) // 2], gap[len(gap) // 10], gap[9 * len(gap) // 10], period), flush=True)
