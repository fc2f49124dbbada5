"""Per-layer critical path of k_tower_tc2 from a SZB_TOWER_TRACE csv (scripts/small_batch.py): for every layer the time from the
previous layer's last release to this layer's, split at the traced stamps."""
import csv, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tower_trace.csv"
want = [int(x) for x in sys.argv[2:]] or [4, 64, 148]
rows = [r for r in csv.DictReader(open(path)) if r["boards"] != "boards"]
by = collections.defaultdict(list)
for r in rows:
    by[int(r["boards"])].append({k: int(v) for k, v in r.items()})
for n in want:
    R = by[n]
    L = collections.defaultdict(list)
    for r in R:
        L[r["layer"]].append(r)
    tot = collections.Counter()
    for l in range(1, 40):
        it = L[l]
        prev = max(r["released_ns"] for r in L[l - 1])
        dep = max(r["dep_ns"] for r in it); ops = max(r["operands_ns"] for r in it)
        acc = max(r["acc_ns"] for r in it); rel = max(r["released_ns"] for r in it)
        tot["release->dependency seen"] += dep - prev; tot["dependency->first chunk in smem"] += ops - dep
        tot["first chunk->accumulator done"] += acc - ops; tot["accumulator->released"] += rel - acc; tot["layer"] += rel - prev
    print("boards %d nsplit %d: mean ns over layers 1..39 (last item of each layer): %s" % (n, R[0]["nsplit"], {k: v // 39 for k, v in tot.items()}))
