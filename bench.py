#!/usr/bin/env python
"""Headline benchmark: synaptic events/s (+ simulated-seconds per wall-second) of the Brunel
network on N B200s, next to the reference's single-threaded CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[4], weak-scaled): Brunel at p = 0.02 with in-degree-scaled weights
(SURVEY §8d C5), sized so that every GPU holds ~5e9 synapses: N(G) = 2e6 * sqrt(G/8) neurons.  At
G = 8 this is the named 2M-neuron / 4e10-synapse network; at G = 1 it is its per-GPU share
(707,100 neurons, 5.0e9 synapses), the largest Brunel that is one GPU's part of that run.
A bench "step" is a block of TIME_STEPS_PER_BENCH_STEP = 150 simulation time steps (150 x snn::step(),
dt = 0.1 ms: ten 15-step delivery windows), taken after PREROLL = 300 untimed time steps that bring the
network to its steady-state firing rates — in BOTH arms (the B200 arm and `--impl reference`), so that
`--steps 20 --warmup 5` times 3,000 time steps of the same steady state the reference arm samples.

The line printed by rank 0 follows the driver's contract; see README/DESIGN for the extra keys
(`parity_check`: the raster of a 200,000-neuron network on the same ranks against the compiled
reference; `generation`: the synapse generator on BASELINE configs[1]).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DT = 1e-4
DELAY = 15e-4
P_CONN = 0.02
METRIC = "synaptic_events_per_sec"
UNIT = "events/s"
TIME_STEPS_PER_BENCH_STEP = 150  # ten delivery windows of 15 time steps
PROFILE_EVERY = 4                # the delivery / update / exchange phases are timed (CUDA events) in every 4th window
PREROLL = 300                    # untimed time steps before the warm-up: the E/I populations start firing at step ~110


def neurons_for(gpus: int, per_gpu_scale: float = 1.0) -> int:
    n = 2.0e6 * math.sqrt(gpus / 8.0) * math.sqrt(per_gpu_scale)
    q = 10 * gpus  # population sizes N/2, 4N/10, N/10 stay integral and split evenly
    return int(round(n / q)) * q


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def raster_digest(counts: np.ndarray, ids: np.ndarray):
    """Order-sensitive digests of a raster (counts[steps, npops] int64, ids int32 in (step, pop) order):
    sha256 (fast in C) — bit-exact equality of the two byte strings is what is compared."""
    import hashlib

    return (hashlib.sha256(np.ascontiguousarray(counts, np.int64).tobytes()).hexdigest()[:16],
            hashlib.sha256(np.ascontiguousarray(ids, np.int32).tobytes()).hexdigest()[:16])


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref, the unmodified reference
    compiled from its sources) on a bounded sample of the workload: the same Brunel construction
    at a size one host core finishes in seconds, the same bench-step definition as the B200 arm."""
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import RefShim

    n = args.ref_neurons
    if not RefShim.available("fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    shim = RefShim("fast")
    w_exc, w_inh = np.float32(0.2 / (P_CONN * n)), np.float32(-1.0 / (P_CONN * n))
    per = args.time_steps
    run = shim.brunel_open(n, P_CONN, w_exc, w_inh, DT, DELAY, 1337)
    run.advance(PREROLL)
    for _ in range(args.warmup):
        run.advance(per)
    sec = ev = 0
    for _ in range(args.steps):
        s_, e_, _sp = run.advance(per)
        sec += s_
        ev += e_
    run.close()
    ev_s = ev / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": ev_s, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"brunel-c5-share: Brunel p={P_CONN}, in-degree-scaled weights; CPU sample N={n}; a bench step = "
                               f"{per} time steps, after {PREROLL} untimed pre-roll time steps (same definition as the B200 arm)",
                   "neurons": n, "synapses": int(P_CONN * n * n / 2), "dt": DT, "delay_steps": 15,
                   "time_steps_per_bench_step": per, "preroll": PREROLL},
        "sim_s_per_wall_s": args.steps * per * DT / sec,
        "cpu_baseline": {"value": ev_s, "unit": UNIT, "cores": 1, "kind": "reference",
                         "sample": f"reference build (-O2 -ffast-math) of Brunel N={n}, p={P_CONN}: {args.steps} x {per} timed time steps "
                                   f"in snn::step(); build {run.build_seconds:.2f}s, timed {sec:.2f}s, 1 thread (the reference has no threading)"},
        "e2e": {"value": ev_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def parity_check(args, rank, world, local_rank, dist, torch):
    """The benchmarked code path against the reference: the same Brunel construction at
    N = --parity-neurons (the network the reference arm builds), on the same ranks, for
    --parity-steps time steps; every spike id of every step is compared with the compiled reference
    (IEEE-strict flavour: bit-exact contract; the verbatim-flags flavour is reported
    beside it)."""
    import spice2_b200 as sp  # noqa: F401
    from spice2_b200.samples import brunel_scaled

    n, steps = args.parity_neurons, args.parity_steps
    q = 10 * world
    n = n // q * q
    net, _pops = brunel_scaled(n, P_CONN, dt=DT, delay=DELAY, device=local_rank, rank=rank, world=world)
    net.finalize()
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, net.peer_handle())
        net.set_peers(handles)
    net.raster_enable(True)
    net.step(steps)
    counts, ids = net.raster_read(steps)
    net.sync()
    events = net.stats()["synaptic_events"]
    if world > 1:
        t = torch.tensor([float(events)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        events = int(t.item())
        dist.barrier()
    net.close()
    if rank != 0:
        return None
    d_counts, d_ids = raster_digest(counts, ids)
    out = {"neurons": n, "steps": steps, "ranks": world, "spikes": int(counts.sum()), "synaptic_events": int(events),
           "sha_counts": d_counts, "sha_ids": d_ids, "matches_reference": None}
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle_lib import RefShim

    w_exc, w_inh = np.float32(0.2 / (P_CONN * n)), np.float32(-1.0 / (P_CONN * n))
    for flavour, key in (("strict", "matches_reference"), ("fast", "matches_reference_fast_math_build")):
        if not RefShim.available(flavour):
            out[key] = None
            out.setdefault("note", "oracle/_ref not present")
            continue
        r = RefShim(flavour).brunel(N=n, p=P_CONN, w_exc=w_exc, w_inh=w_inh, dt=DT, delay=DELAY, seed=1337, steps=steps)
        same = bool(np.array_equal(r["counts"], counts) and np.array_equal(r["ids"], ids))
        out[key] = same
        if flavour == "strict":
            rc, ri = raster_digest(r["counts"], r["ids"])
            out["reference_sha_counts"], out["reference_sha_ids"] = rc, ri
            # events: the reference tallies a spike's events when it is delivered (delay - 1 steps after it was emitted)
            out["reference_flavour"] = "reference sources, -fno-fast-math -ffp-contract=off (bit-exact contract)"
    return out


def generation_check(local_rank):
    """Synapse generation on BASELINE configs[1]: fixed_probability(0.1), seed {1337}: 1e4 x 1e4 against the
    golden hash of the compiled reference's adjacency, 1e5 x 1e5 (999,991,208 edges) timed on the device."""
    import spice2_b200 as sp

    out = {}
    gold = json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())
    g4 = next((c for c in gold.get("fixed_probability", []) if c["src"] == 10000 and c["dst"] == 10000 and abs(c["p"] - 0.1) < 1e-12
               and c.get("increments", 0) == 0), None)
    r = sp.generate_fixed_probability(10000, 10000, 0.1, (1337,), device=local_rank)
    out["edges_1e4"] = r["edges"]
    if g4 is not None:  # FNV-1a64 of the compiled reference's arrays (tests/golden/make_golden.py)
        out["golden_hash_ok"] = bool(r["edges"] == g4["edges"] and sp.fnv1a64(r["offsets"]) == g4["fnv_offsets"]
                                     and sp.fnv1a64(r["neighbors"]) == g4["fnv_neighbors"])
    best = None
    for _ in range(3):  # the generation has host round trips per chunk of the stream: a busy host shows (69 .. 200 ms seen)
        r5 = sp.generate_fixed_probability(100000, 100000, 0.1, (1337,), device=local_rank, copy=False)
        if best is None or r5["total_ms"] < best["total_ms"]:
            best = r5
    out.update({"edges": best["edges"], "edges_expected": 999991208, "device_ms": best["total_ms"],
                "edges_per_s": best["edges"] / (best["total_ms"] * 1e-3),
                "write_roofline_frac": best["edges"] * 4.0 / (best["total_ms"] * 1e-3) / 1e9 / measured_peaks()[0],
                "config": "fixed_probability(0.1) 1e5 x 1e5, seed {1337} (BASELINE configs[1]); best of 3, CUDA events"})
    out["golden_hash_ok"] = bool(out.get("golden_hash_ok", True) and best["edges"] == 999991208)
    # the counter-based generator (north star: "counter-based RNG and geometric skip sampling ... bounded by write bandwidth"):
    # another matrix of the same distribution, same shape, same seed
    fast = None
    for _ in range(2):
        f5 = sp.generate_fixed_probability(100000, 100000, 0.1, (1337,), device=local_rank, copy=False, fast=True)
        if fast is None or f5["total_ms"] < fast["total_ms"]:
            fast = f5
    out["counter_based"] = {"edges": fast["edges"], "device_ms": fast["total_ms"], "edges_per_s": fast["edges"] / (fast["total_ms"] * 1e-3),
                            "write_roofline_frac": fast["edges"] * 4.0 / (fast["total_ms"] * 1e-3) / 1e9 / measured_peaks()[0],
                            "mean_degree_error_sigmas": (fast["edges"] - 1e9) / (1e10 * 0.1 * 0.9) ** 0.5}
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--time-steps", type=int, default=TIME_STEPS_PER_BENCH_STEP, help="simulation time steps per bench step (both arms)")
    ap.add_argument("--neurons", type=int, default=0, help="override the network size (default: weak-scaled C5 share)")
    ap.add_argument("--ref-neurons", type=int, default=200000)
    ap.add_argument("--cpu-baseline-steps", type=int, default=3000)
    ap.add_argument("--parity-neurons", type=int, default=200000)
    ap.add_argument("--parity-steps", type=int, default=300)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-generation", action="store_true")
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1 and args.time_steps >= 1

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the first communicator comes up; rank 0's stdout carries
        # ONE JSON line, so the communicator is created with fd 1 pointing at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    import spice2_b200 as sp  # noqa: F401
    from spice2_b200.samples import brunel_scaled

    TS = args.time_steps
    n = args.neurons or neurons_for(world)
    t_build0 = time.time()
    stream = torch.cuda.Stream()
    net, (P, E, I) = brunel_scaled(n, P_CONN, dt=DT, delay=DELAY, device=local_rank, rank=rank, world=world)
    net.set_stream(stream.cuda_stream)
    net.finalize()
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, net.peer_handle())
        net.set_peers(handles)
    net.sync()
    build_s = time.time() - t_build0
    synapses_local = sum(net.connection_edges(c) for c in range(6))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(v, op=None):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
        return t.item()

    MAX = dist.ReduceOp.MAX if world > 1 else None

    # ---- device-timed run: inputs resident in HBM -------------------------------------------------
    with torch.cuda.stream(stream):
        net.step(PREROLL)                      # untimed: to the steady-state firing rates
        for _ in range(args.warmup):           # W warm-up bench steps
            net.step(TS)
        barrier()
        st0 = net.stats()
        w0 = net.windows_run()
        net.profile_enable(True, every=PROFILE_EVERY)
        net.profile_read()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):            # K timed bench steps
            net.step(TS)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clk = clocks.stop() if rank == 0 else None
        prof = net.profile_read()
        net.profile_enable(False)
        st1 = net.stats()
        windows_total = net.windows_run() - w0
    time_steps = args.steps * TS
    events_local = st1["synaptic_events"] - st0["synaptic_events"]
    spikes_local = st1["spikes_delivered"] - st0["spikes_delivered"]
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    ms_max = allreduce(ms, MAX)
    events_total = allreduce(events_local)
    synapses_total = allreduce(synapses_local)
    deliver_ms_max = allreduce(prof["deliver_ms"], MAX)
    exchange_ms_max = allreduce(prof["exchange_ms"], MAX)

    # ---- end-to-end through the public API: spikes of every step come back to the host -----------
    e2e = None
    if not args.no_e2e:
        with torch.cuda.stream(stream):
            net.raster_enable(True)
            barrier()
            s0 = net.stats()
            t0 = time.perf_counter()
            d2h = 0
            done = issued = 0
            pending = []
            # one readout per bench step; the host drains bench step k from the page-locked sink while the
            # device runs bench steps k + 1 and k + 2
            while done < args.steps:
                while issued < args.steps and len(pending) < 3:
                    net.step(TS)
                    pending.append(TS)
                    issued += 1
                k = pending.pop(0)
                counts, ids = net.raster_read(k)  # every spike id of these time steps, in host memory
                d2h += counts.nbytes + ids.nbytes
                done += 1
            barrier()
            wall = time.perf_counter() - t0
            s1 = net.stats()
            net.raster_enable(False)
        wall_max = allreduce(wall, MAX)
        ev2 = allreduce(s1["synaptic_events"] - s0["synaptic_events"])
        e2e = {"value": ev2 / wall_max, "unit": UNIT, "h2d_bytes_per_step": 20 * TS, "d2h_bytes_per_step": d2h / args.steps,
               "sim_s_per_wall_s": time_steps * DT / wall_max, "readout_batches": args.steps, "batches_in_flight": 3,
               "note": "per time step the host sends dt + the step's 128-bit stream seed (kernel arguments) and receives every spike id, "
                       "sorted per (step, population) as neuron_population::spikes() returns them; one readout per bench step "
                       f"({TS} time steps), overlapped with the next bench steps on the device"}
    net.close()

    # ---- parity of the benchmarked path against the reference, on the same ranks ---------------------
    parity = None
    if not args.no_parity:
        parity = parity_check(args, rank, world, local_rank, dist, torch)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (spike delivery) ------------------------------------------
    peak, peak_src = measured_peaks()
    # 4 B/event + 16 B offsets + 4 B id per spike.  The delivery launches are timed with CUDA events inside the
    # timed region, every PROFILE_EVERY-th window (four timed events per window cost ~3 % of the step): the average launch
    # duration of the sampled launches against the average algorithmic bytes of a launch
    alg_bytes = 4.0 * events_local + 20.0 * spikes_local
    alg_per_launch = alg_bytes / max(1, windows_total)
    achieved = alg_per_launch * prof["windows"] / (prof["deliver_ms"] * 1e-3) / 1e9 if prof["deliver_ms"] > 0 else 0.0
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "traffic_r02.json"
    if tf.exists() and world == 1 and not args.neurons:
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel in this workload's steady state, from the
        # committed `ncu --set full` capture; that launch's own algorithmic bytes are given beside it (bench.py cannot run
        # under ncu: a number printed under a profiler is never a bench value)
        try:
            j = json.loads(tf.read_text())
            traffic = j.get("deliver_dram_bytes_per_launch")
            traffic_src = {"file": "profiles/traffic_r02.json", "algorithmic_bytes_of_captured_launch": j.get("algorithmic_bytes_of_that_launch"),
                           "ratio_in_capture": j.get("ratio")}
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "deliver_units (1 launch per 15-step window, all 6 connections)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_per_launch,
                "deliver_us_per_launch": prof["deliver_ms"] * 1e3 / max(1, prof["windows"]),
                "deliver_ms_total": prof["deliver_ms"], "update_ms_total": prof["update_ms"],
                "exchange_ms_total": prof["exchange_ms"], "exchange_ms_max_over_ranks": exchange_ms_max, "windows": prof["windows"],
                "windows_timed_of": [prof["windows"], windows_total],
                "timing": f"CUDA events inside the timed region around the phases of every {PROFILE_EVERY}th window; the *_ms_total are sums over those windows",
                "deliver_share_of_step": deliver_ms_max * windows_total / max(1, prof["windows"]) / ms_max}

    # ---- synapse generation (BASELINE configs[1]) ----------------------------------------------------
    generation = None
    if not args.no_generation:
        try:
            generation = generation_check(local_rank)
        except Exception as e:  # noqa: BLE001 - reported in the line, the headline numbers stand
            generation = {"error": str(e)}

    # ---- CPU baseline: the compiled reference on a bounded sample ----------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, str(ROOT / "tests"))
        from oracle_lib import RefShim

        nref = args.ref_neurons
        if RefShim.available("fast"):
            run = RefShim("fast").brunel_open(nref, P_CONN, np.float32(0.2 / (P_CONN * nref)), np.float32(-1.0 / (P_CONN * nref)),
                                              DT, DELAY, 1337)
            run.advance(PREROLL)  # untimed pre-roll to the steady-state rates
            sec, ev, _sp = run.advance(args.cpu_baseline_steps)
            run.close()
            cpu = {"value": ev / sec, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"compiled reference (its own flags), Brunel N={nref} p={P_CONN}, {args.cpu_baseline_steps} time steps after a "
                             f"{PREROLL}-step pre-roll: build {run.build_seconds:.2f}s, snn::step() loop {sec:.2f}s on 1 of {os.cpu_count()} host cores "
                             f"(the reference is single-threaded)",
                   "sim_s_per_wall_s": args.cpu_baseline_steps * DT / sec}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref not present"}

    line = {
        "metric": METRIC, "value": events_total / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"brunel-c5-share: Brunel p={P_CONN}, in-degree-scaled weights, {n} neurons on {world} GPU(s) "
                               f"(BASELINE configs[4] weak-scaled: ~5e9 synapses per GPU); a bench step = {TS} time steps, "
                               f"after {PREROLL} untimed pre-roll time steps",
                   "neurons": n, "synapses": int(synapses_total), "dt": DT, "delay_steps": 15, "window_steps": 15,
                   "time_steps_per_bench_step": TS, "preroll": PREROLL, "timed_time_steps": time_steps,
                   "l2": "per-window delivery streams ~0.8 GB of synapse rows per GPU (> 126 MB L2)", "mode": "deterministic"},
        "sim_s_per_wall_s": time_steps * DT / (ms_max * 1e-3),
        "ms_per_time_step": ms_max / time_steps,
        "build_s": build_s,
        "gpu_launches": int(launches),
        "clocks": clk,
        "e2e": e2e,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity_check": parity,
        "generation": generation,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
