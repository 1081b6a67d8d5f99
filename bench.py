#!/usr/bin/env python3
"""bench.py -- cell-timesteps/s of the SUEWS per-grid-cell timestep physics on B200.

Workload (BASELINE.json configs[2]): synthetic 1 048 576 cells per GPU, BaseRun cell with per-cell
surface fractions / location / morphology / forcing offsets, default scheme set
(NetRadiationMethod 3, OHM, LUMPS first guess, StabilityMethod 3, snow off), forcing = the reference's
2004 hourly series disaggregated to 5 min (the whole leap year, 105 408 rows, resident on the device).
One bench "step" advances every cell by one day (288 timesteps) in ONE kernel launch.  The K timed days are spread
evenly over the year (day 366*i/K of 2004: with the driver's --steps 20 every 18th day), so that the figure is the
year's mix of winter / summer, day / night physics and not that of one month; the state is carried from launch to
launch, only the forcing rows jump.

  value  whole-job cell-timesteps/s with forcing + state resident in HBM (CUDA-event time of the
         kernel launches, max over ranks)
  e2e    the same through the public host API with host buffers: per step the day's forcing goes
         host -> device and the per-cell running sums of the 7 parity fluxes come back device -> host
  --impl reference   the CPU oracle (a C port of the Fortran: no Fortran compiler exists in this
         image) on all host cores, one process per core, bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "cell_timesteps_per_s"
UNIT = "cell-timesteps/s"
PARITY_COLS = [5, 6, 7, 8, 9, 16, 71]  # QN QF QS QH QE RO SWE (0-based in the 82-column line)
N_STATE, N_CELL, N_SUM = 208, 46, 82


def load_inputs(ndays, day0=0):
    import fixtures as fx
    base = fx.state_init(2005)
    from suews_b200.forcing import disaggregate_met, read_met
    hourly = read_met(os.path.join(fx.GOLDEN, "forcing_2004_60min.txt"))
    nh = min(hourly.shape[0], (day0 + ndays) * 24 + 48)
    met = disaggregate_met(hourly[:nh], last_block=(nh == hourly.shape[0]), next_row=hourly[min(nh, hourly.shape[0] - 1)])
    # the 5-min series starts at 00:05; a day = 288 rows
    return base, met[day0 * 288:(day0 + ndays) * 288]


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------ CPU arm (oracle)
def _cpu_worker(args):
    worker, budget_s, max_cells, scheme, days = args
    import fixtures as fx
    import oracle_lib as ol
    from suews_b200 import synth
    base, met_year = load_inputs(366)
    d = synth.base_with_scheme(base, scheme)
    site = synth.cell_site_arrays(worker * max_cells, max_cells, d)
    ts5 = ts5_year = None
    if synth.SCHEMES[scheme]["storageheatmethod"] in (4, 14):
        d.update({k.lower(): v for k, v in synth.estm_example(d["surfacearea"][0]).items()})
        ts5_year = synth.estm_ts_forcing(met_year)
    t0 = time.perf_counter()
    done = 0
    for c in range(max_cells):
        day = days[c % len(days)]                      # one day per cell, cycling through the bench's days of the year
        met = met_year[day * 288:(day + 1) * 288]
        if ts5_year is not None:
            ts5 = ts5_year[day * 288:(day + 1) * 288]
        dd = dict(d)
        for k in ("sfr", "lat", "lng", "bldgH", "EveTreeH", "DecTreeH", "Z"):
            dd[k.lower()] = np.atleast_1d(site[k][c])
        cell = ol.Cell(dd)
        m = met.copy()
        m[:, 11] += site["dTemp_C"][c]
        m[:, 10] = np.minimum(100.0, np.maximum(np.float64(np.float32(0.01)), m[:, 10] + site["dRH"][c]))
        if ts5 is not None:
            cell.estm_start(ts5[0])
        cell.run(m, want=(), ts5=ts5)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0


def cpu_arm(budget_s, scheme, cores=None, days=(0,)):
    import multiprocessing as mp
    import oracle_lib as ol
    ol.build()
    ol.lib()   # mapped in this process too (the forked workers inherit it): the oracle is what this arm runs, visibly
    cores = cores or os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(w, budget_s, 4096, scheme, list(days)) for w in range(cores)])
        wall = time.perf_counter() - t0
    cells = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return cells * 288 / slowest, cores, cells, slowest, wall


def year_days(K, W, ndays_year=366):
    """forcing day of each launch: W warm-up days ahead of the first timed day, then K days spread over the year"""
    timed = [min(ndays_year - 1, (i * ndays_year) // max(K, 1)) for i in range(K)]
    warm = [(timed[0] + ndays_year - W + i) % ndays_year for i in range(W)] if K else list(range(W))
    return warm, timed


def run_scheme(SuewsBatch, synth, base, met, ncell, rank, local_rank, scheme, days_warm, days_timed, OUT, barrier_extra=None):
    """device-resident timing of one scheme set on the given forcing days; returns (seconds, trips per step, fatal cells)"""
    use_estm = synth.SCHEMES[scheme]["storageheatmethod"] in (4, 14)
    b = SuewsBatch(ncell, 1, device=local_rank)
    ts5 = synth.estm_ts_forcing(met) if use_estm else None
    synth.build_batch(b, base, cell0=rank * ncell, scheme=scheme, ts_row0=ts5[0] if use_estm else None)
    b.upload()
    b.set_forcing(met)
    if use_estm:
        b.set_estm_forcing(ts5)
    for d in days_warm:
        b.advance(d * 288, 288, OUT, sync=False)
    b.device_sync()
    ms = []
    for d in days_timed:
        b.advance(d * 288, 288, OUT, sync=False)
        ms.append(b.last_ms())
    b.device_sync()
    fatal, _warn, iters = b.diag()
    nl = len(days_warm) + len(days_timed)
    out = (sum(ms) * 1e-3, float(iters.sum()) / (ncell * 288.0 * nl), int((fatal != 0).sum()))
    b.close()
    return out


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells-per-gpu", type=int, default=1048576)
    ap.add_argument("--scheme", default="default")
    ap.add_argument("--cpu-budget", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flux-only", action="store_true", help="skip the side measurement with SUEWS_OUT_ALL_STAGES")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the end-to-end region (default: --steps)")
    ap.add_argument("--consecutive-days", type=int, default=-1,
                    help="time K consecutive days starting at this day of 2004 instead of K days spread over the year")
    ap.add_argument("--scheme-steps", type=int, default=4,
                    help="timed days (spread over the year) of each extra scheme set in the `schemes` object; 0 = skip it")
    ap.add_argument("--schemes", default="baserun,snow,estm,netrad13,anohm", help="scheme sets of the `schemes` object (configs[3])")
    ap.add_argument("--total-cells", type=int, default=0,
                    help="strong scaling (BASELINE configs[4]): this many cells in total, split in contiguous blocks over the GPUs; "
                         "use with --consecutive-days 0 --steps 30 for the 16.8 M cells x 30 days case")
    ap.add_argument("--single-process", type=int, default=0,
                    help="drive this many GPUs from ONE process through the library's multi-device handle (suews_b200_multi_*) "
                         "instead of one process per GPU; the end-to-end region then includes the final gather to the first device")
    ap.add_argument("--sums", default="parity7", choices=["parity7", "all82", "hourly"],
                    help="on-device output reduction: P0 = running sums of the 7 parity fluxes (SURVEY 8d), of all 82 columns, "
                         "or P1 = the hourly-aggregated 87-column SUEWS block of every cell (SUEWS_Output_Agg in the kernel)")
    a = ap.parse_args()
    K, W = a.steps, max(a.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ngpu = a.single_process if a.single_process > 0 else world     # GPUs the job runs on
    if a.total_cells > 0:       # strong scaling: contiguous blocks of the global cell index
        per = a.total_cells // ngpu
        a.cells_per_gpu = per + (1 if rank < a.total_cells % ngpu else 0)
        cell0_rank = rank * per + min(rank, a.total_cells % ngpu)
    else:
        cell0_rank = rank * a.cells_per_gpu
    if a.consecutive_days >= 0:
        days_warm = [(a.consecutive_days + i) % 366 for i in range(W)]
        days_timed = [(a.consecutive_days + W + i) % 366 for i in range(K)]
    else:
        days_warm, days_timed = year_days(K, W)
    workload = ("synthetic %s x 1 year (2004) of 5-min forcing sampled as %d one-day launches of 288 steps (days-of-year %s; "
                "warm-up days %s), %s scheme set (%s), shared forcing + per-cell dT/dRH") % (
        ("%d cells in total over %d GPUs" % (a.total_cells, ngpu)) if a.total_cells > 0 else ("%d cells/GPU" % a.cells_per_gpu),
        K, ",".join(str(d + 1) for d in days_timed), ",".join(str(d + 1) for d in days_warm), a.scheme,
        "see suews_b200/synth.py SCHEMES")

    if a.impl == "reference":
        if rank != 0:
            return
        rates, info = [], None
        for i in range(W + K):
            r, cores, cells, slowest, wall = cpu_arm(max(2.0, a.cpu_budget / 2), a.scheme, days=days_timed)
            if i >= W:
                rates.append((cells * 288, slowest))
            info = (cores, cells)
        tot_cs = sum(x[0] for x in rates)
        tot_t = sum(x[1] for x in rates)
        val = tot_cs / tot_t
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * tot_t / max(K, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": "%d cells x 288 steps per step over %d host processes" % (info[1], info[0])},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": info[0], "kind": "port",
                             "sample": "%d cells x 288 steps per step; CPU oracle (C port of the Fortran; no Fortran compiler in the image)" % info[1]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from suews_b200 import OUT_AGG, OUT_SUM7, OUT_SUM82, SuewsBatch, synth
    from suews_b200 import _lib
    from suews_b200.batch import all_fields
    global N_STATE, N_CELL
    N_STATE = sum(n for _nm, kind, n in all_fields() if kind == 3)
    N_CELL = sum(n for _nm, kind, n in all_fields() if kind == 2)

    OUT = {"parity7": OUT_SUM7, "all82": OUT_SUM82, "hourly": OUT_AGG}[a.sums]
    n_sum = 7 if a.sums == "parity7" else N_SUM
    use_estm = synth.SCHEMES[a.scheme]["storageheatmethod"] in (4, 14)
    base, met = load_inputs(366)   # the whole year 2004 at 5 min: 105 408 rows, 16.9 MB of step records on the device
    sch = synth.SCHEMES[a.scheme]
    workload = workload.replace("see suews_b200/synth.py SCHEMES", "snowUse %d, NetRadiationMethod %d, StorageHeatMethod %d, StabilityMethod %d, EmissionsMethod %d" % (
        sch["snowuse"], sch["netradiationmethod"], sch["storageheatmethod"], sch["stabilitymethod"], sch["emissionsmethod"]))
    ncell = a.cells_per_gpu
    if a.single_process > 0:     # one process, several GPUs: the library's own partition + final gather
        from suews_b200 import SuewsMultiBatch
        ncell = a.total_cells if a.total_cells > 0 else a.cells_per_gpu * ngpu
        b = SuewsMultiBatch(ncell, 1, devices=list(range(ngpu)))
    else:
        b = SuewsBatch(ncell, 1, device=local_rank)
    ts5 = synth.estm_ts_forcing(met) if use_estm else None   # synthetic Ts forcing = air temperature
    synth.build_batch(b, base, cell0=cell0_rank, scheme=a.scheme, ts_row0=ts5[0] if use_estm else None)
    b.upload()

    def barrier():
        b.device_sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident timing
    b.set_forcing(met)
    if use_estm:
        b.set_estm_forcing(ts5)
    for d in days_warm:
        b.advance(d * 288, 288, OUT, sync=False)
    barrier()
    _f0, _w0, iters_before = b.diag()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    ms = []
    for d in days_timed:
        b.advance(d * 288, 288, OUT, sync=False)
        ms.append(b.last_ms())   # CUDA events recorded on the library's launch stream
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    b.sync()   # raises if any cell hit a fatal SUEWS error
    dev_s = sum(ms) * 1e-3
    if dist is not None:
        import torch
        t = torch.tensor([dev_s, wall], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, wall = float(t[0]), float(t[1])
    cells_job = float(a.total_cells) if a.total_cells > 0 else (float(ncell) if a.single_process > 0 else float(world) * ncell)
    total_cs = cells_job * 288 * K
    value = total_cs / dev_s

    _f0, _w0, iters_after = b.diag()
    trips_per_step = float(iters_after.sum() - iters_before.sum()) / (ncell * 288.0 * max(K, 1))

    # ---------------- side measurement (not the headline): policy P0 forms only what its seven sums can observe, so the
    # roughness-sublayer diagnostics T2/q2/U10/RH2 are not evaluated; SUEWS_OUT_ALL_STAGES runs them anyway
    all_stages = None
    if a.sums == "parity7" and not a.no_flux_only:
        from suews_b200.batch import OUT_ALL_STAGES
        ms2 = []
        sub = days_timed[:: max(1, len(days_timed) // 4)][:4]
        for d in sub:
            b.advance(d * 288, 288, OUT | OUT_ALL_STAGES, sync=False)
            ms2.append(b.last_ms())
        barrier()
        fo_s = sum(ms2) * 1e-3
        if dist is not None:
            import torch
            t = torch.tensor([fo_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fo_s = float(t[0])
        all_stages = {"value": cells_job * 288 * len(sub) / fo_s, "unit": UNIT,
                      "note": "SUEWS_OUT_ALL_STAGES: every stage of SUEWS_cal_Main, including the RSL diagnostics T2/q2/U10/RH2 that the seven "
                              "flux sums of policy P0 cannot observe; identical sums and state; days-of-year %s re-run from the advanced "
                              "state; NOT the headline value" % ",".join(str(d + 1) for d in sub)}

    # ---------------- end-to-end through the host API (host buffers in, host results out)
    e2e_t0 = None
    gather_ms = []
    cols = PARITY_COLS
    Ke = a.e2e_steps if a.e2e_steps > 0 else K
    e2e_days = days_warm + [days_timed[i % max(K, 1)] for i in range(Ke)]   # the same days again, forcing from the host
    for i, d0 in enumerate(e2e_days):
        if i == W:
            barrier()
            e2e_t0 = time.perf_counter()
        day = met[d0 * 288:(d0 + 1) * 288]
        b.set_forcing(day)                      # host -> device (pinned staging) + ephemeris kernel
        if use_estm:
            b.set_estm_forcing(ts5[d0 * 288:(d0 + 1) * 288])
        b.advance(0, 288, OUT, sync=False)
        if a.sums == "hourly":                  # device -> host: the last hour's QN..QE maps of all cells
            res = np.stack([b.agg_column(23, 5 + c) for c in cols], axis=1)
        else:
            if a.single_process > 0:            # the final gather: device -> first device (peer copies) -> host
                res, gms = b.gather_sums(cols)
                gather_ms.append(gms)
            else:
                res = b.sums_cols(cols)         # device -> host (pinned staging), waits for the kernel
    barrier()
    e2e_s = time.perf_counter() - e2e_t0
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_val = cells_job * 288 * Ke / e2e_s
    checksum = float(np.sum(res[:, 0]))
    nonfinite_cells = int((~np.isfinite(res).all(axis=1)).sum())
    fatal, warn, iters = b.diag()
    b.close()

    # ---------------- BASELINE configs[3]: the other scheme sets on the same cells, days spread over the year (N = 1 only)
    schemes = None
    if ngpu == 1 and a.scheme_steps > 0 and a.sums == "parity7":
        schemes = {}
        _w, sdays = year_days(a.scheme_steps, 0)
        sdays = [min(365, d + 9) for d in sdays]   # (days 10, 101, 192, 284 for 4 steps: mid-winter to autumn)
        for name in [x for x in a.schemes.split(",") if x]:
            sec, trips, nfatal = run_scheme(SuewsBatch, synth, base, met, ncell, rank, local_rank, name, [sdays[0] - 1], sdays, OUT)
            sc = synth.SCHEMES[name]
            schemes[name] = {"value": ncell * 288.0 * len(sdays) / sec, "unit": UNIT, "steps": len(sdays), "warmup": 1,
                             "days_of_year": [d + 1 for d in sdays], "tsurf_iterations_per_step": trips, "fatal_cells": nfatal,
                             "switches": "snowUse %d, NetRadiationMethod %d, StorageHeatMethod %d" % (
                                 sc["snowuse"], sc["netradiationmethod"], sc["storageheatmethod"])}

    if rank != 0:
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    # algorithmic bytes per launch: state read+write, cell params read, sums read+write, per cell; forcing shared
    ncell_gpu = ncell // ngpu if a.single_process > 0 else ncell      # cells of one device = of one kernel launch
    bytes_per_launch = ncell_gpu * 8.0 * (2 * N_STATE + N_CELL + 2 * n_sum) + 288 * 160.0
    if a.sums == "hourly":   # P1: no sums in/out; 24 aggregated rows x 87 columns written per cell
        bytes_per_launch = ncell_gpu * 8.0 * (2 * N_STATE + N_CELL + 24 * 87) + 288 * 160.0
    launch_s = dev_s / K
    ach_gbs = bytes_per_launch / launch_s / 1e9
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json")))
    except Exception:
        pass
    import ctypes as C
    tf = C.c_double()
    fp64_peak = None
    if _lib.lib().suews_b200_fp64_peak(local_rank, C.byref(tf)) == 0:
        fp64_peak = tf.value
    # executed FP64 flops (DFMA = 2) from the ncu captures, averaged over four points of the year (tools/run_captures.sh)
    fp64 = {"peak_tflops_measured_dfma": fp64_peak, "flops_per_cell_step_executed": prof.get("fp64_flops_per_cell_step"),
            "source": prof.get("source")}
    if fp64_peak and prof.get("fp64_flops_per_cell_step"):
        fp64["achieved_tflops"] = prof["fp64_flops_per_cell_step"] * ncell_gpu * 288 / launch_s / 1e12
        fp64["frac"] = fp64["achieved_tflops"] / fp64_peak
    # the bound that matters: ALGORITHMIC flops of the hoisted minimal algorithm, counted by the oracle's op
    # counter with the published libm weights (tools/oracle_opcount.py -> profiles/oracle_opcount_<scheme>.json), x the
    # measured rate, against the DFMA peak measured on this device
    roofline_fp64 = None
    try:
        oc = json.load(open(os.path.join(ROOT, "profiles", "oracle_opcount_%s.json" % a.scheme)))
        falg = float(oc["hoisted"]["weighted_flops_per_cell_step"])
        ach = falg * value / float(ngpu) / 1e12
        roofline_fp64 = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": (ach / fp64_peak) if fp64_peak else None, "flops_per_cell_step_algorithmic": falg,
                         "flops_per_cell_step_as_written": float(oc["literal"]["weighted_flops_per_cell_step"]),
                         "peak_source": "DFMA throughput measured on this device (suews_b200_fp64_peak)",
                         "source": "oracle op counter, hoisted minimal algorithm, days-of-year %s (profiles/oracle_opcount_%s.json); "
                                   "weights add/mul/div/sqrt 1, exp 20, log 25, pow 60, sin/cos 30, atan 35" % (oc["days_of_year"], a.scheme)}
    except Exception:
        pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * dev_s / K, "higher_is_better": True, "scaling": "strong" if a.total_cells > 0 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "cells_per_gpu": ncell_gpu, "timesteps_per_step": 288,
                   "processes": ("one process drives %d GPUs through suews_b200_multi_* (final gather to the first device inside e2e)" % ngpu)
                   if a.single_process > 0 else "one process per GPU", "output_policy": {
                       "parity7": "P0: per-cell running sums of the 7 parity fluxes QN,QF,QS,QH,QE,RO,SWE",
                       "all82": "P0+: per-cell running sums of all 82 SUEWS columns",
                       "hourly": "P1: hourly-aggregated 87-column SUEWS block of every cell, aggregated in the kernel"}[a.sums],
                   "l2": "inputs (state+params %.0f MB/GPU) larger than L2" % (ncell_gpu * 8.0 * (N_STATE + N_CELL + N_SUM) / 1e6),
                   "tsurf_iterations_per_step": trips_per_step, "wall_s_timed": wall,
                   "checksum_QN": checksum, "fatal_cells": int((fatal != 0).sum()), "nonfinite_sum_cells": nonfinite_cells},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 288 * 160, "d2h_bytes_per_step": ncell * len(cols) * 8},
        "gpu_launches": K * (ngpu if a.single_process > 0 else 1),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                     "traffic": (prof["dram_bytes_per_cell_step"] * ncell_gpu * 288.0 if "dram_bytes_per_cell_step" in prof else None), "peak_source": peak_src,
                     "note": "kernel is FP64-pipe / latency bound, not HBM bound (SURVEY 8d); see roofline_fp64"},
        "fp64": fp64,
    }
    if roofline_fp64:
        line["roofline_fp64"] = roofline_fp64
    if gather_ms:
        line["e2e"]["final_gather_ms_per_step"] = float(np.mean(gather_ms[W:])) if len(gather_ms) > W else float(np.mean(gather_ms))
    if all_stages:
        line["all_stages"] = all_stages
    if schemes:
        line["schemes"] = schemes
    if ngpu == 1 and not a.no_cpu_baseline:
        r, cores, cells, slowest, _ = cpu_arm(a.cpu_budget, a.scheme, days=days_timed)
        line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d cells x 288 steps (one day each, cycling through the bench's days of the year) of the same synthetic cells, one process per core (%.1f s)" % (cells, slowest)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
