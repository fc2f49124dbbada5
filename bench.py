#!/usr/bin/env python
"""bench.py -- the headline measurement: MCTS simulations/s (and self-play moves/s) at 800 simulations per move
over >= 1024 concurrent chess games per B200, beside the reference's CPU path on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

A "step" is one pass of the hot path over one batch: one self-play ply for every game of the batch =
`num_searches` MCTS simulations per game (select -> move generation / encoding -> network -> expand / backup),
then a move sampled from the visit counts and pushed (szb_selfplay_ply; reference: sim.py:46-76).

  value   : whole-job simulations/s with the games resident in HBM (device-timed, CUDA events on the library's
            stream, max over ranks);
  e2e     : the same metric through the host-buffer API: every step uploads the batch's positions from pinned
            host memory (szb_games_set), searches, and reads the visit counts + child masks back to pinned host
            memory (szb_search) -- the call the MCTS0.search facade makes;
  roofline: the dominant kernel -- k_tower_tc2, the whole 41-layer tower as ONE persistent tcgen05 launch per 512 boards -- timed
            on the device inside the timed region (the kernel stamps %globaltimer at its first CTA's start and its last CTA's end;
            the two cohorts' launches overlap on two streams, which CUDA events cannot bracket), against the measured sustained
            bf16 peak of MEASURED_PEAKS.json;
  extras  : (same JSON line) config c3 (4096 Chess960 games) as a rate, a bounded config-c4 iteration (500 games sharded over
            the ranks, strong scaling) next to the weak-scaling headline, and the fine-tuning step of the same loop on the
            library's trainer (f1_trainer: positions/s at batch 128, torch eager on the same GPU beside it);
  cpu_baseline / --impl reference: the restated reference search (oracle/ref_path.py: the reference's Python
            control flow + torch CPU fp32 network, batch 1, all host threads) on a bounded sample of the workload.

oracle/ is used here only as that CPU baseline; the measured product path is libszb200.so (no fallback).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mcts_simulations_per_sec"
UNIT = "simulations/s"
FLOP_PER_EVAL = 2914845184                    # 2 x MACs of one policyNN forward
WORKLOADS = {
    # BASELINE.json configs[2] / configs[3]
    "c2": dict(games=1024, sims=800, chess960=False,
               name="c2: 1024 concurrent vanilla-chess games x 800 sims/move, bf16 network, C=2, learning=True"),
    "c3": dict(games=4096, sims=800, chess960=True,
               name="c3: 4096 concurrent Chess960 games x 800 sims/move, bf16 network, C=2, learning=True"),
    # BASELINE.json configs[4]: one whole self-play iteration, its games split over the ranks (strong scaling); not a bench line
    # of the driver (a step is a whole iteration of complete games: minutes), run by hand: --workload c4 [--gpus N]
    "c4": dict(games=500, sims=800, chess960=True,
               name="c4: full self-play iteration, num_selfPlay_iterations=500 Chess960 games x 800 sims/move to the end of every "
                    "game, sharded over the ranks, NCCL weight broadcast first"),
}
C_PUCT = 2.0
SEED = 0
MAX_PREFIX = 40                                # S2 "random-played" start positions: U{0..40} random legal plies


def common_config(wl, weights):
    """the workload description BOTH arms print (identical keys and values, so the driver's same_config check holds)"""
    return {"workload": wl["name"], "games_per_gpu": wl["games"], "num_searches": wl["sims"], "C": C_PUCT, "learning": True,
            "chess960": wl["chess960"], "positions": "S2 random-played: U{0..40} random legal plies from the start, seed 0",
            "weights": weights}


def runtime_flat(model, dev):
    from sigma_zero_b200 import runtime
    return runtime.flat_weights(model, dev)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="B200_PROFILING.md fallback")


def traffic_per_launch(kind, boards):
    """dram read+write bytes of one launch of the timed kernel from the committed ncu --set full capture
    (profiles/traffic.json), if a capture of that kernel at that batch exists"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        for e in json.load(open(p)).get("captures", []):
            # (a launch of the timed region may be a few boards short of the captured batch: finished games take no row)
            if int(e["conv_kind"]) == int(kind) and abs(int(e["boards"]) - int(boards)) <= 0.02 * int(e["boards"]):
                return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------
# workload: S2 random-played start positions (same generator for both arms: ascending policy-index order)
# ---------------------------------------------------------------------------------------------------
def prefix_rng(g):
    return np.random.default_rng([SEED, int(g)])


def start_id_of(g, chess960):
    return (SEED + int(g)) % 960 if chess960 else -1


def play_prefixes_gpu(eng, game_ids, chess960):
    """random legal plies through the engine (szb_legal_moves / szb_games_push); returns plies played per game"""
    n = len(game_ids)
    eng.reset([start_id_of(g, chess960) for g in game_ids])
    rngs = [prefix_rng(g) for g in game_ids]
    want = np.array([int(r.integers(0, MAX_PREFIX + 1)) for r in rngs])
    played = np.zeros(n, dtype=np.int64)
    for ply in range(MAX_PREFIX):
        idx, cnt = eng.legal_moves()
        who, mv = [], []
        for g in range(n):
            if ply < want[g] and cnt[g] > 0:
                k = int(rngs[g].integers(0, int(cnt[g])))
                who.append(g)
                mv.append(int(idx[g, k]))
        if not who:
            break
        eng.push(who, mv)
        played[who] += 1
    return played


def oracle_prefix_game(g, chess960):
    """the same prefix for one game through the oracle (reference arm / cpu baseline)"""
    from oracle import ref_path
    sid = start_id_of(g, chess960)
    game = ref_path.RefGame(chess960=chess960, start_id=sid if chess960 else None)
    rng = prefix_rng(g)
    want = int(rng.integers(0, MAX_PREFIX + 1))
    for _ in range(want):
        b = game.board
        pairs = sorted((ref_path.move_to_index(m, b.turn), m) for m in b.legal_moves)
        if not pairs:
            break
        k = int(rng.integers(0, len(pairs)))
        game.move_piece(pairs[k][1])
    return game


def seeded_model():
    """policyNN default init under torch.manual_seed(0), eval mode (the real weights are a git-LFS pointer in the
    reference); a user-supplied supervised_model_best.pt (> 1 KB) is used instead when present"""
    import torch
    from sigma_zero_b200.network import policyNN
    torch.manual_seed(0)
    model = policyNN({}).eval()
    for cand in (os.path.join(ROOT, "supervised_model_best.pt"), "supervised_model_best.pt"):
        if os.path.exists(cand) and os.path.getsize(cand) > 1024:
            model.load_state_dict(torch.load(cand, map_location="cpu"))
            return model, "file:" + cand
    return model, "random-init torch.manual_seed(0)"


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc = device, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU reference (oracle) timing
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(chess960, sims_per_step, steps, warmup, game=0):
    """restated reference search on the host: one game, one move decision of `sims_per_step` simulations per step
    (batch-1 torch CPU fp32 forward per simulation, mcts.py:72-75).  Returns (sims/s, seconds per step, threads)."""
    import torch
    from oracle import ref_path
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = ref_path.build_policy_nn().eval()
    ev = ref_path.torch_evaluator(model)
    g = oracle_prefix_game(game, chess960)
    if g.board.outcome() is not None:
        g = oracle_prefix_game(game + 1, chess960)
    for _ in range(warmup):
        ref_path.search(g, max(8, sims_per_step // 8), C_PUCT, ev, learning=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_path.search(g, sims_per_step, C_PUCT, ev, learning=True)
    dt = time.perf_counter() - t0
    return sims_per_step * steps / dt, dt / steps, threads


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sims = args.ref_sims
    rate, sec_step, threads = cpu_reference_run(wl["chess960"], sims, args.steps, args.warmup)
    sample = "1 game (workload game 0, random-played prefix) x 1 move decision x %d of the %d simulations per step" % (sims, wl["sims"])
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": common_config(wl, "random-init torch.manual_seed(0)"),
        "arm": "reference CPU path: python control flow + torch CPU fp32 batch-1 network over the oracle chess stand-in, all host threads",
        "host_cores": threads,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from sigma_zero_b200 import _lib
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)

    G, S = wl["games"], wl["sims"]
    eng = Engine(max_games=G, max_searches=S, device=local)
    model, weights = seeded_model()

    # weights: rank 0's flat fp32 buffer broadcast over NCCL (the only collective of the path) and folded + packed ON THE GPU straight
    # from the broadcast buffer (szb_net_load_device): nothing crosses the host
    from sigma_zero_b200 import runtime
    keys, numels, flat = runtime.flat_weights(model, dev)
    bcast_ms = None
    if world > 1:
        dist.broadcast(flat, src=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(flat, src=0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    eng.load_flat_device(keys, numels, flat)                      # first load: allocations, tensor maps
    torch.cuda.synchronize()
    t_load = time.perf_counter()
    eng.load_flat_device(keys, numels, flat)                      # steady state (what every later iteration pays): in-place refill
    load_ms = (time.perf_counter() - t_load) * 1e3
    digest = eng.net_checksum()
    digests_equal = None
    if world > 1:
        lo_hi = torch.tensor([digest & 0xFFFFFFFF, digest >> 32], dtype=torch.int64, device=dev)
        mine = lo_hi.clone()
        dist.broadcast(lo_hi, src=0)
        same = torch.tensor([int(bool((lo_hi == mine).all()))], dtype=torch.int64, device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        digests_equal = bool(int(same.item()))

    # weak scaling: every rank owns its own block of G games (global game ids rank*G .. rank*G+G-1)
    game_ids = list(range(rank * G, rank * G + G))
    play_prefixes_gpu(eng, game_ids, wl["chess960"])

    ext = torch.cuda.ExternalStream(eng.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident run (value) ---------------------------------------------------------------
    for w in range(args.warmup):
        eng.selfplay_ply(S, C_PUCT, True, EVAL_NET_BF16, seed=SEED, sample=True)
    st0 = eng.stats()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    eng.tower_spans(True)              # k_tower_tc2 stamps its own start / end (%globaltimer): per-launch durations of the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    sims_done = moves_done = 0
    for k in range(args.steps):
        moves, _ = eng.selfplay_ply(S, C_PUCT, True, EVAL_NET_BF16, seed=SEED, sample=True)
        live = int((moves >= 0).sum())
        sims_done += live * S
        moves_done += live
    e1.record(ext)
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    spans = eng.tower_spans(False)
    st1 = eng.stats()
    # measurement pass (not part of `value`): one more step with the library's profiling on, which serialises the two
    # cohorts on the context's stream and brackets every phase and the network's tower launch with CUDA events
    eng.set_profiling(True)
    eng.selfplay_ply(S, C_PUCT, True, EVAL_NET_BF16, seed=SEED, sample=True)
    pt = eng.phase_times()
    eng.set_profiling(False)
    total_sims = sum_over_ranks(sims_done)
    total_moves = sum_over_ranks(moves_done)
    launches = int(sum_over_ranks(st1["kernel_launches"] - st0["kernel_launches"]))
    evals = int(sum_over_ranks(st1["evaluations"] - st0["evaluations"]))
    value = total_sims / (ms * 1e-3)

    # ---- end-to-end run through host buffers (e2e) --------------------------------------------------
    pos_bytes = G * 112
    pin_pos = torch.empty(pos_bytes, dtype=torch.uint8, pin_memory=True)
    pin_vis = torch.empty((G, 4672), dtype=torch.int32, pin_memory=True)
    pin_child = torch.empty((G, 73), dtype=torch.int64, pin_memory=True)
    cur = eng.positions()
    import ctypes
    ctypes.memmove(pin_pos.data_ptr(), ctypes.addressof(cur), pos_bytes)
    np_pos = pin_pos.numpy()
    np_vis = pin_vis.numpy().view(np.uint32)
    np_child = pin_child.numpy().view(np.uint64)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        eng.set_positions_buffer(np_pos, G)                                   # H2D: G x szb_pos
        eng.search(S, C_PUCT, True, EVAL_NET_BF16, out=(np_vis, np_child, None))    # D2H: visits + child masks
        return int(np_vis.sum(dtype=np.int64))

    e2e_step()                                                                # warm-up
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(ext)
    visit_sum = 0
    for k in range(e2e_steps):
        visit_sum += e2e_step()
    f1.record(ext)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), wall_ms))                # host-side copies count: take the longer clock
    live_games = int((np_vis.sum(axis=1) > 0).sum())
    e2e_value = sum_over_ranks(live_games * S * e2e_steps) / (e2e_ms * 1e-3)
    assert visit_sum == live_games * (S - 1) * e2e_steps, "visit counts must sum to num_searches-1 per live game"

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    pk = peaks()
    roof = None
    kernel_name = {2: "k_tower_tc2: stem + 38 tower 3x3 convolutions + policy 1x1 in one persistent CTA-pair launch (tcgen05 cta_group::2 implicit GEMM)",
                   1: "k_tower_tc2 on one 3x3 256->256 tower convolution (CTA-pair tcgen05 implicit GEMM)",
                   0: "k_conv_tc<256,0> on one 3x3 256->256 tower convolution (single-CTA tcgen05 implicit GEMM)"}[pt["conv_kind"]]
    profiled = None
    if pt["conv_launches"] > 0:
        conv_ms = pt["conv_ms"] / pt["conv_launches"]
        profiled = {"ms_per_launch": conv_ms, "achieved": pt["conv_flop"] / (conv_ms * 1e-3) / 1e12, "launches": pt["conv_launches"],
                    "boards_per_launch": pt["conv_boards"],
                    "note": "CUDA events around every launch of one extra step run serialised (one cohort) right after the timed region; "
                            "≈ 10 % idle time between launches lets the power-capped clock recover, so this is a little faster than the timed region"}
    if spans and spans["launches"] > 0:
        # the timed region itself: every tower launch of the timed steps, timed on the device by the kernel (first CTA start -> last CTA
        # end); the two cohorts' launches overlap by a few us at their edges, so busy time can exceed wall time slightly
        ms_launch = spans["busy_ns"] / spans["launches"] * 1e-6
        # Overlap-corrected: the two cohorts' launches overlap at their edges (and, since round 2, run beside the other cohort's tree
        # kernel), so the SUM of the per-launch spans counts the overlapped intervals twice.  The time that bounds the tower's work is
        # the time during which at least one tower launch ran: the sum of spans where they do not overlap, else the window they cover.
        exclusive_ns = min(spans["busy_ns"], spans["wall_ns"]) if spans["wall_ns"] else spans["busy_ns"]
        achieved = spans["flop"] / (exclusive_ns * 1e-9) / 1e12
        achieved_sum_of_spans = spans["flop"] / (spans["busy_ns"] * 1e-9) / 1e12
        boards = spans["boards"] / spans["launches"]
        roof = {"bound": "tensor", "kernel": kernel_name, "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"], "frac_of_burst": achieved / pk["bf16_burst"],
                "peak_kind": "sustained, " + pk["source"],
                "how": "every tower launch of the TIMED REGION stamped on the device by the kernel itself (%globaltimer, first CTA start to last CTA "
                       "end; CUDA events cannot bracket kernels of two overlapping streams); achieved = their algorithmic FLOP / the time at least "
                       "one of them was running (overlap-corrected: min(sum of spans, window they cover))",
                "achieved_sum_of_spans": achieved_sum_of_spans,
                "traffic": traffic_per_launch(pt["conv_kind"], round(boards)), "ms_per_launch": ms_launch,
                "launches_timed": spans["launches"], "flop_per_launch": spans["flop"] / spans["launches"], "boards_per_launch": boards,
                # (share of the window the recorded launches span: the library keeps the first 8192 launches of the timed region)
                "tower_busy_share_of_timed_region": spans["busy_ns"] / spans["wall_ns"] if spans["wall_ns"] else None,
                # the SM clock the tower actually ran at: clock64 cycles / %globaltimer ns of CTA 0 of every launch (nvidia-smi samples at
                # 200 ms cannot see the dips of a power-capped step); utilisation = achieved / (148 SMs x 8192 FLOP/clk x that clock)
                "sm_mhz_inside_launches": 1e3 * spans["sm_cycles"] / spans["sm_ns"] if spans.get("sm_ns") else None,
                "tensor_pipe_utilisation_at_that_clock": (achieved * 1e12 / (148 * 8192 * 1e9 * spans["sm_cycles"] / spans["sm_ns"])) if spans.get("sm_ns") else None,
                "frac_timed_region_lower_bound": FLOP_PER_EVAL * evals / (ms * 1e-3) / 1e12 / world / pk["bf16_sustained"],
                "profiled_step": profiled}
    elif profiled:
        roof = {"bound": "tensor", "kernel": kernel_name, "achieved": profiled["achieved"], "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": profiled["achieved"] / pk["bf16_sustained"], "frac_of_burst": profiled["achieved"] / pk["bf16_burst"],
                "peak_kind": "sustained, " + pk["source"], "traffic": traffic_per_launch(pt["conv_kind"], pt["conv_boards"]),
                "ms_per_launch": profiled["ms_per_launch"], "launches_timed": profiled["launches"], "flop_per_launch": pt["conv_flop"],
                "boards_per_launch": pt["conv_boards"], "profiled_step": profiled}
    steps_total = max(1, pt["steps"])
    phases = {k: pt[k] / steps_total for k in ("select_ms", "expand_ms", "eval_ms", "finish_ms")}
    net_tflops = (FLOP_PER_EVAL * G) / (phases["eval_ms"] * 1e-3) / 1e12 if phases["eval_ms"] > 0 else None
    # the HBM-class kernels of the step against the copy peak, algorithmic bytes of SURVEY.md 8d / DESIGN.md 3.2-3.3: latency-bound
    # at a search's batch size by construction (a few hundred bytes per tree behind 2-4 dependent round trips)
    def hbm(bytes_, ms):
        gbs = bytes_ / (ms * 1e-3) / 1e9 if ms > 0 else None
        return {"algorithmic_bytes": int(bytes_), "ms": ms, "achieved_gbs": gbs, "frac_of_hbm_peak": (gbs / pk["hbm_gbs"]) if gbs else None}
    hbm_kernels = {
        "k_select": hbm(16 * pt["select_edges"] + 16 * pt["select_levels"], pt["select_ms"]),
        "k_expand": hbm(steps_total * G * (96 + 96 + 7 * 96 + 952 + 584), pt["expand_ms"]),
        "k_finish": hbm(24 * pt["backup_levels"] + 24 * pt["edges_written"], pt["finish_ms"]),
        "note": "totals over the profiled step; these kernels move ~2 MB per launch at 1024 trees and are latency-bound (hidden under the "
                "other cohort's tower in the timed region); their bandwidth-regime numbers (65,536 trees, bulk perft) are in profiles/r01f_*",
    }

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n = args.cpu_sims
        rate, sec, threads = cpu_reference_run(wl["chess960"], n, 1, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "1 game (workload game 0) x 1 move decision x %d simulations, torch CPU fp32 batch-1 network (%.1f s)" % (n, sec)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": common_config(wl, weights),
            "arm": {"parallelism": "games sharded, %d x network replica" % world,
                    "pipelining": ("2 cohorts of %d games on two streams (tree kernel of one runs under the other's tower)" % (G // 2))
                                  if G >= 768 else "one cohort (the library splits batches of >= 768 running games into two)",
                    "step": "k_tree_step (children + backup of the previous leaves, PUCT descent, move generation, planes, network input rows) "
                            "+ k_tower_tc2 (41 layers, priors of the legal moves and value head in the last epilogue): 2 launches per 512 boards",
                    "l2": "no flush: per-step working set (3 x %d MB activations + 46 MB weights + tree arena) exceeds the 126 MB L2"
                          % (min(G, 1024) * 100 * 256 * 2 // 2 ** 20)},
            "host_cores": os.cpu_count(),
            "moves_per_sec": total_moves / (ms * 1e-3), "evals_per_sec": evals / (ms * 1e-3),
            "network_tflops_in_step": net_tflops,
            # lower bound of the tower's rate inside the timed region itself (two cohorts pipelined): every network evaluation of the
            # timed steps at its algorithmic FLOP over the whole elapsed time, as if nothing but the tower ran
            "network_tflops_timed_region": FLOP_PER_EVAL * evals / (ms * 1e-3) / 1e12,
            "phase_ms_per_simulation_step": phases,
            "phase_note": "phases (and roofline.profiled_step) come from one extra profiled step (single cohort, CUDA events on the library stream) run right after the timed region; roofline.achieved is timed on the device inside the timed region",
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pos_bytes,
                    "d2h_bytes_per_step": G * 4672 * 4 + G * 73 * 8, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "szb_games_set(host positions) + szb_search(host visit/child buffers), pinned memory"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "hbm_kernels": hbm_kernels, "cpu_baseline": cpu,
            "weights_broadcast_ms": bcast_ms, "weights_load_ms": load_ms, "weights_digest_equal_on_all_ranks": digests_equal,
            "weights_note": "flat fp32 state_dict (91 MB) broadcast over NCCL, then BatchNorm fold + bf16/fp32 packing by the library's kernels "
                            "from the device buffer (szb_net_load_device, host wall clock incl. the stream sync); digest = szb_net_checksum",
            "lib": _lib.LIB_PATH.replace(ROOT + os.sep, ""),
        }
    else:
        line = None
    eng.close()
    return line


def run_extras(args):
    """Configs c3 and c4 next to the c2 headline, in the same JSON line (`extras`): c3 = 4096 concurrent Chess960 games x 800 sims as a
    rate (one timed ply, weak scaling like the headline); c4 = one self-play iteration of 500 Chess960 games x 800 sims SHARDED over
    the ranks (strong scaling: fixed total work), bounded to --extras-c4-plies plies per game so that the default run stays short."""
    import torch
    import torch.distributed as dist
    from sigma_zero_b200.engine import EVAL_NET_BF16, Engine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    out = {}
    model, weights = seeded_model()

    def reduce(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())
    MAX = dist.ReduceOp.MAX if world > 1 else None
    SUM = dist.ReduceOp.SUM if world > 1 else None

    # ---- c3 -------------------------------------------------------------------------------------------
    wl = WORKLOADS["c3"]
    G, S = wl["games"], wl["sims"]
    eng = Engine(max_games=G, max_searches=S, device=local)
    eng.load_flat_device(*runtime_flat(model, dev))
    play_prefixes_gpu(eng, list(range(rank * G, rank * G + G)), True)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    eng.selfplay_ply(S, C_PUCT, True, EVAL_NET_BF16, seed=SEED, sample=True)           # warm-up ply
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    moves, _ = eng.selfplay_ply(S, C_PUCT, True, EVAL_NET_BF16, seed=SEED, sample=True)
    e1.record(ext)
    torch.cuda.synchronize()
    ms = reduce(e0.elapsed_time(e1), MAX)
    live = reduce(int((moves >= 0).sum()), SUM)
    eng.close()
    out["c3"] = {"workload": wl["name"], "metric": METRIC, "value": live * S / (ms * 1e-3), "unit": UNIT, "scaling": "weak",
                 "games_per_gpu": G, "running_games": int(live), "steps": 1, "warmup": 1, "ms_per_step": ms, "n_gpus": world}
    # ---- c4, bounded ------------------------------------------------------------------------------------
    c4 = c4_measure(args, dict(WORKLOADS["c4"]), args.extras_c4_plies, model, weights)
    if c4 is not None:
        out["c4_bounded"] = {k: c4[k] for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "moves_per_sec", "positions_recorded",
                                                "rank_seconds", "weights_broadcast_ms")}
        out["c4_bounded"]["workload"] = c4["config"]["workload"] + " [bounded: %d plies per game]" % args.extras_c4_plies
        out["c4_bounded"]["games_per_gpu"] = c4["config"]["games_per_gpu"]
    # ---- f1: the fine-tuning step of the same loop (train_RL.py:77-154) on the library's trainer, next to torch eager -----------
    if rank == 0:
        try:
            out["f1_trainer"] = trainer_measure(model, dev, local)
        except Exception as e:                                     # an extra must never take the headline down
            out["f1_trainer"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out if rank == 0 else None


def trainer_measure(model, dev, local, batch=128, steps=20, warmup=3):
    """positions/s of one optimiser step (forward + backward + Adam) at the reference's batch size (train_RL.py:175: 128) on synthetic
    records, through sigma_zero_b200.trainer.Trainer (szb_train_step: row numbers from the host every step), timed with CUDA events on
    the library's stream; beside it the same step in torch eager on the same GPU (the reference's own trainer code path, convolutions
    in torch's default TF32 mode), inputs already on the device"""
    import numpy as np
    import torch
    from sigma_zero_b200 import records as records_mod
    from sigma_zero_b200.engine import Engine
    from sigma_zero_b200.trainer import Trainer
    rng = np.random.default_rng(SEED)
    n = 8 * batch
    bits = rng.random((n, 119, 64)) < 0.12
    idx, prob, off = [], [], [0]
    for _ in range(n):
        k = int(rng.integers(5, 45))
        idx.extend(np.sort(rng.choice(4672, size=k, replace=False)).tolist())
        p = rng.random(k).astype(np.float32)
        prob.extend((p / p.sum()).tolist())
        off.append(len(idx))
    rec = {"states": np.packbits(bits, axis=-1, bitorder="little").view("<u8").reshape(n, 119).astype(np.uint64),
           "pi_index": np.array(idx, np.uint16), "pi_prob": np.array(prob, np.float32), "pi_off": np.array(off, np.int64),
           "z": rng.integers(-1, 2, n).astype(np.int8)}
    batches = [rng.permutation(n)[:batch].astype(np.int32) for _ in range(8)]
    eng = Engine(max_games=2, max_searches=8, device=local)
    tr = Trainer(eng, model, batch_size=batch)
    tr.set_records(rec)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    first = tr.step(batches[0])
    for i in range(1, warmup):
        tr.step(batches[i % 8], want_losses=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for i in range(steps):
        last = tr.step(batches[i % 8], want_losses=(i == steps - 1))
    e1.record(ext)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tr.close()
    eng.close()
    # torch eager, same step
    ref = type(model)({}).to(dev).train()
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4, weight_decay=1e-4)
    x = torch.from_numpy(records_mod.unpack_states(rec, batches[0])).to(device=dev, dtype=torch.float32)
    pi = torch.from_numpy(records_mod.dense_policy(rec, batches[0])).to(dev)
    z = torch.from_numpy(rec["z"][batches[0]].astype(np.float32)).to(dev)

    def torch_step():
        opt.zero_grad(set_to_none=True)
        p, v = ref.forward_torch(x)
        (torch.nn.functional.mse_loss(v.squeeze(-1), z) + torch.nn.functional.cross_entropy(p, pi)).backward()
        opt.step()
    for _ in range(warmup):
        torch_step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        torch_step()
    t1.record()
    torch.cuda.synchronize()
    ms_torch = t0.elapsed_time(t1) / 10
    flop = 3 * 2 * 64 * batch * (256 * (119 * 9 + 38 * 2304 + 256) + 73 * 256)       # forward + dgrad + wgrad of every convolution (no dgrad for the stem: ~1 % less)
    return {"metric": "trained_positions_per_sec", "value": batch / (ms * 1e-3), "unit": "positions/s", "batch": batch, "ms_per_step": ms,
            "steps": steps, "warmup": warmup, "tflops": flop / (ms * 1e-3) / 1e12, "dtype": "bf16 operands, fp32 accumulate / master weights / Adam",
            "loss_first_step": list(first), "loss_last_step": list(last),
            "torch_eager_same_gpu": {"ms_per_step": ms_torch, "value": batch / (ms_torch * 1e-3), "mode": "fp32 parameters, TF32 convolutions (torch default)"},
            "speedup_vs_torch_eager": ms_torch / ms, "data": "synthetic records, resident on the GPU; row numbers sent from the host every step"}


def c4_measure(args, wl, max_plies, model, weights):
    """One whole self-play iteration through the product API (train_RL.selfplay_iteration's sharding + sim.selfplay_records):
    every game of this rank's block is played to its end (or to max_plies) on the GPU, records land in packed host arrays.
    Returns the JSON line on rank 0 (None elsewhere); the process group, if any, must be up."""
    import torch
    import torch.distributed as dist
    from sigma_zero_b200 import runtime
    from sigma_zero_b200.sim import selfplay_records
    from sigma_zero_b200.train_RL import broadcast_weights, shard_of

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    sargs = {"C": C_PUCT, "num_searches": wl["sims"], "num_selfPlay_iterations": wl["games"], "chess960": wl["chess960"],
             "leaves_per_tree": args.leaves_per_tree}
    lo, hi = shard_of(wl["games"], rank, world)
    # the parameters live on the GPU, as in train_RL.main(train_device="cuda"): the broadcast buffer is then built, sent and folded
    # without leaving the device
    model.to(dev)
    # warm-up: two plies of this rank's block (engine creation, weight upload, first launches) and one broadcast (NCCL communicator)
    selfplay_records(model, sargs, hi - lo, c960=wl["chess960"], seed=SEED, max_plies=2, game_id_base=lo)
    eng = runtime.get_engine()
    if world > 1:
        broadcast_weights(model, 0, dev, eng)
        torch.cuda.synchronize()
    st0 = eng.stats()
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    bcast_ms = None
    if world > 1:
        tb = time.perf_counter()
        broadcast_weights(model, 0, dev, eng)                 # NCCL broadcast + GPU-side fold straight from the broadcast buffer
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - tb) * 1e3
    rec, counters = selfplay_records(model, sargs, hi - lo, c960=wl["chess960"], seed=SEED, max_plies=max_plies, game_id_base=lo)
    torch.cuda.synchronize()
    my_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    wall_s = time.perf_counter() - t0
    clocks = sampler.stop()
    st1 = eng.stats()

    def reduce(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())
    R = dist.ReduceOp if world > 1 else None
    total_s = reduce(wall_s, R.MAX if R else None)
    sims = reduce(counters["simulations"], R.SUM if R else None)
    moves = reduce(len(rec["z"]), R.SUM if R else None)
    evals = reduce(st1["evaluations"] - st0["evaluations"], R.SUM if R else None)
    launches = reduce(st1["kernel_launches"] - st0["kernel_launches"], R.SUM if R else None)
    unfinished = reduce(int((rec["result"] == 2).sum()), R.SUM if R else None)
    longest = reduce(counters["plies"], R.MAX if R else None)
    slowest_rank_s = reduce(my_s, R.MAX if R else None)
    fastest_rank_s = -reduce(-my_s, R.MAX if R else None)
    rec_bytes = reduce(sum(v.nbytes for v in rec.values()), R.SUM if R else None)
    if rank == 0:
        line = {
            "metric": METRIC, "value": sims / total_s, "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 0,
            "ms_per_step": total_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": wl["name"], "games": wl["games"], "games_per_gpu": (wl["games"] + world - 1) // world,
                       "num_searches": wl["sims"], "C": C_PUCT, "learning": True, "chess960": True, "weights": weights,
                       "start": "Chess960 ids drawn per game from (seed, global game id)", "max_plies": max_plies,
                       "leaves_per_tree": args.leaves_per_tree,
                       "search": "the reference's algorithm (one simulation of a tree at a time)" if args.leaves_per_tree <= 1 else
                                 "NON-PARITY multi-leaf mode: %d simulations of a tree in flight per step, virtual loss" % args.leaves_per_tree,
                       "timed": "weight broadcast + every ply of every game (search, sample, push, packed record to host) until the "
                                "last game of the slowest rank ends; host wall clock between barriers"},
            "moves_per_sec": moves / total_s, "evals_per_sec": evals / total_s, "positions_recorded": int(moves),
            "games_unfinished_at_max_plies": int(unfinished), "longest_game_plies": int(longest),
            "rank_seconds": {"slowest": slowest_rank_s, "fastest": fastest_rank_s}, "record_bytes": int(rec_bytes),
            "e2e": {"value": sims / total_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(rec_bytes),
                    "note": "the timed region is the end-to-end API call: records arrive in host memory every ply"},
            "gpu_launches": int(launches), "clocks": clocks, "weights_broadcast_ms": bcast_ms,
        }
        return line
    return None


def init_dist():
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--games", type=int, default=None, help="override games per GPU (parity / debugging; not a bench line)")
    ap.add_argument("--sims", type=int, default=None, help="override simulations per move (debugging)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-sims", type=int, default=800, help="--impl reference: simulations per step (bounded sample)")
    ap.add_argument("--cpu-sims", type=int, default=2000, help="cpu_baseline: simulations in the bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--leaves-per-tree", type=int, default=1,
                    help="--workload c4 only: opt-in multi-leaf search with virtual loss (not the reference's algorithm; reported separately)")
    ap.add_argument("--max-plies", type=int, default=None, help="--workload c4: cut games after this many plies (default: play every game out)")
    ap.add_argument("--no-extras", action="store_true", help="skip the c3 rate, the bounded c4 iteration and the trainer step that ride on the default (c2) line")
    ap.add_argument("--extras-c4-plies", type=int, default=24, help="plies per game of the bounded c4 iteration in `extras`")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.games:
        wl["games"] = args.games
        wl["name"] += " [games overridden: %d]" % args.games
    if args.sims:
        wl["sims"] = args.sims
        wl["name"] += " [sims overridden: %d]" % args.sims
    if args.impl == "reference":
        run_reference(args, wl)
        return
    dist = init_dist()
    rank = int(os.environ.get("RANK", "0"))
    if args.workload == "c4":
        model, weights = seeded_model()
        line = c4_measure(args, wl, args.max_plies, model, weights)
    else:
        line = run_ours(args, wl)
        if args.workload == "c2" and not args.no_extras and not args.games and not args.sims:
            try:
                extras = run_extras(args)
            except Exception as e:                     # the headline line must survive a failure of the riders
                extras = {"error": "%s: %s" % (type(e).__name__, e)}
            if line is not None:
                line["extras"] = extras
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
