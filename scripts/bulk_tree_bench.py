"""Bulk (bandwidth-regime) measurements of the tree / move-generation kernels: many trees with the integer-hash
evaluator (no network), per-phase CUDA-event times from the library, algorithmic bytes from its counters
(SURVEY.md 8d per-unit figures).  Also bulk perft (legal move generation over millions of positions).
Usage: python scripts/bulk_tree_bench.py [games=65536] [sims=48]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from sigma_zero_b200 import _lib
from sigma_zero_b200.engine import Engine, EVAL_HASH
from tests import util

G = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
S = int(sys.argv[2]) if len(sys.argv) > 2 else 48
HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
eng = Engine(max_games=G, max_searches=S)
eng.reset([-1 if g % 2 == 0 else g % 960 for g in range(G)])
eng.search(8, 2.0, True, EVAL_HASH, want_visits=False, want_children=False)      # warm-up
eng.set_profiling(True)
eng.search(S, 2.0, True, EVAL_HASH, want_visits=False, want_children=False)
pt = eng.phase_times()
eng.set_profiling(False)
steps = pt["steps"]
sel_bytes = 16 * pt["select_edges"] + 16 * pt["select_levels"]
fin_bytes = 24 * pt["backup_levels"] + 20 * pt["edges_written"] + 4 * pt["edges_written"] + 8 * 80 * G * steps
exp_bytes = (96 + 96 + 7 * 96 + 952 + 584) * G * steps
out = {"games": G, "sims": S, "hbm_peak_gbs": HBM}
for name, ms, b in (("select", pt["select_ms"], sel_bytes), ("expand", pt["expand_ms"], exp_bytes), ("finish", pt["finish_ms"], fin_bytes)):
    gbs = b / (ms * 1e-3) / 1e9
    out[name] = {"ms_per_step": ms / steps, "algorithmic_bytes_per_step": b / steps, "GBps": gbs, "frac_of_hbm_peak": gbs / HBM}
    print("%-7s %8.3f ms/step  %8.1f MB/step  %8.1f GB/s  %.3f of measured HBM peak" % (name, ms / steps, b / steps / 1e6, gbs, gbs / HBM))
print("sims/s (tree + movegen only, hash evaluator): %.0f" % (G * S / ((pt["select_ms"] + pt["expand_ms"] + pt["eval_ms"] + pt["finish_ms"]) * 1e-3)))
eng.close()
# bulk perft: last-level counting kernel = legal move generation over n positions (96 B read each)
eng = Engine(max_games=4, max_searches=4)
for fen_name, pos, depth in (("startpos", util.wire_pos(util.oracle_game(False, -1).board, _lib), 6),):
    nodes, ms, npos = eng.perft(pos, depth, timed=True)
    gbs = 96 * npos / (ms * 1e-3) / 1e9
    out["perft_" + fen_name] = {"depth": depth, "nodes": nodes, "positions_last_level": npos, "ms": ms, "Mpos_per_s": npos / ms / 1e3, "GBps": gbs, "frac_of_hbm_peak": gbs / HBM}
    print("perft(%d) %s = %d; movegen over %d positions in %.3f ms: %.1f M positions/s, %.1f GB/s (%.3f of HBM peak)" % (depth, fen_name, nodes, npos, ms, npos / ms / 1e3, gbs, gbs / HBM))
eng.close()
json.dump(out, open("gpurun_out/bulk_tree_bench.json", "w"), indent=1)
