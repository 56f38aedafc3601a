#!/usr/bin/env python
"""Headline benchmark: synaptic events/s (+ simulated-seconds per wall-second) of the Brunel
network on N B200s, next to the reference's single-threaded CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[4], weak-scaled): Brunel at p = 0.02 with in-degree-scaled weights
(SURVEY §8d C5), sized so that every GPU holds ~5e9 synapses: N(G) = 2e6 * sqrt(G/8) neurons.  At
G = 8 this is the named 2M-neuron / 4e10-synapse network; at G = 1 it is its per-GPU share
(707,100 neurons, 5.0e9 synapses), the largest Brunel that is one GPU's part of that run.
A "step" is one simulation time step (snn::step(), dt = 0.1 ms).

The line printed by rank 0 follows the driver's contract; see README/DESIGN for the extra keys.
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
def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (oracle/_ref, the unmodified reference
    compiled from its sources) on a bounded sample of the workload: the same Brunel construction
    at a size one host core finishes in seconds."""
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
    total = args.warmup + args.steps
    # A bench "step" of this arm is a bounded sample: `per` consecutive time steps of the reference's
    # own snn::step() on the same Brunel construction at a size one host core handles (the reference
    # is single-threaded).  The network first runs PREROLL untimed time steps so that the sample is
    # taken at the steady-state firing rates, like the GPU arm's.
    per = args.ref_steps or max(1, min(200, 6000 // max(1, total)))
    PREROLL = 300
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
    ms_per_step = sec / (args.steps * per) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": ev_s, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"brunel-c5-share: Brunel p={P_CONN}, in-degree-scaled weights; CPU sample N={n}, "
                               f"{per} time step(s) per bench step after {PREROLL} untimed pre-roll steps",
                   "neurons": n, "synapses": int(P_CONN * n * n / 2), "dt": DT, "delay_steps": 15},
        "sim_s_per_wall_s": args.steps * per * DT / sec,
        "cpu_baseline": {"value": ev_s, "unit": UNIT, "cores": 1, "kind": "reference",
                         "sample": f"reference build (-O2 -ffast-math) of Brunel N={n}, p={P_CONN}: {args.steps} x {per} timed time steps "
                                   f"in snn::step(); build {run.build_seconds:.2f}s, timed {sec:.2f}s, 1 thread (the reference has no threading)"},
        "e2e": {"value": ev_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--neurons", type=int, default=0, help="override the network size (default: weak-scaled C5 share)")
    ap.add_argument("--ref-neurons", type=int, default=200000)
    ap.add_argument("--ref-steps", type=int, default=0, help="time steps per bench step in the reference arm (0: sized from --steps)")
    ap.add_argument("--cpu-baseline-steps", type=int, default=3000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

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

    import spice2_b200 as sp
    from spice2_b200.samples import brunel_scaled

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

    # ---- device-timed run: inputs resident in HBM -------------------------------------------------
    with torch.cuda.stream(stream):
        net.step(args.warmup)
        barrier()
        st0 = net.stats()
        net.profile_enable(True)
        net.profile_read()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        net.step(args.steps)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clk = clocks.stop() if rank == 0 else None
        prof = net.profile_read()
        net.profile_enable(False)
        st1 = net.stats()
    events_local = st1["synaptic_events"] - st0["synaptic_events"]
    spikes_local = st1["spikes_delivered"] - st0["spikes_delivered"]
    launches = st1["kernel_launches"] - st0["kernel_launches"]

    def allreduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return t.item()

    ms_max = allreduce(ms, dist.ReduceOp.MAX if world > 1 else None)
    events_total = allreduce(events_local, dist.ReduceOp.SUM if world > 1 else None)
    synapses_total = allreduce(synapses_local, dist.ReduceOp.SUM if world > 1 else None)
    deliver_ms_max = allreduce(prof["deliver_ms"], dist.ReduceOp.MAX if world > 1 else None)

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
            batch = 150  # ten windows per readout
            pending = []
            # the host drains batch k from the page-locked sink while the device runs batch k + 1
            while done < args.steps:
                while issued < args.steps and len(pending) < 2:
                    k = min(batch, args.steps - issued)
                    net.step(k)
                    pending.append(k)
                    issued += k
                k = pending.pop(0)
                counts, ids = net.raster_read(k)  # every spike id of these steps, in host memory
                d2h += counts.nbytes + ids.nbytes
                done += k
            barrier()
            wall = time.perf_counter() - t0
            s1 = net.stats()
            net.raster_enable(False)
        wall_max = allreduce(wall, dist.ReduceOp.MAX if world > 1 else None)
        ev2 = allreduce(s1["synaptic_events"] - s0["synaptic_events"], dist.ReduceOp.SUM if world > 1 else None)
        e2e = {"value": ev2 / wall_max, "unit": UNIT, "h2d_bytes_per_step": 20, "d2h_bytes_per_step": d2h / args.steps,
               "sim_s_per_wall_s": args.steps * DT / wall_max,
               "note": "per step the host sends dt + the step's 128-bit stream seed (kernel arguments) and receives every spike id, "
                       "sorted per (step, population) as neuron_population::spikes() returns them, in batches of 150 steps"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (spike delivery) ------------------------------------------
    peak, peak_src = measured_peaks()
    alg_bytes = 4.0 * events_local + 20.0 * spikes_local  # 4 B/event + 16 B offsets + 4 B id per spike
    achieved = alg_bytes / (prof["deliver_ms"] * 1e-3) / 1e9 if prof["deliver_ms"] > 0 else 0.0
    traffic = None
    tf = ROOT / "profiles" / "traffic_r01.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("deliver_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "deliver_tiles (1 launch per 15-step window, all 6 connections)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(1, prof["windows"]),
                "deliver_ms_total": prof["deliver_ms"], "update_ms_total": prof["update_ms"],
                "exchange_ms_total": prof["exchange_ms"], "windows": prof["windows"],
                "deliver_share_of_step": deliver_ms_max / ms_max}

    # ---- CPU baseline: the compiled reference on a bounded sample ----------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        sys.path.insert(0, str(ROOT / "tests"))
        from oracle_lib import RefShim

        nref = args.ref_neurons
        if RefShim.available("fast"):
            run = RefShim("fast").brunel_open(nref, P_CONN, np.float32(0.2 / (P_CONN * nref)), np.float32(-1.0 / (P_CONN * nref)),
                                              DT, DELAY, 1337)
            run.advance(300)  # untimed pre-roll to the steady-state rates
            sec, ev, _sp = run.advance(args.cpu_baseline_steps)
            run.close()
            cpu = {"value": ev / sec, "unit": UNIT, "cores": 1, "kind": "reference",
                   "sample": f"compiled reference (its own flags), Brunel N={nref} p={P_CONN}, {args.cpu_baseline_steps} time steps after a "
                             f"300-step pre-roll: build {run.build_seconds:.2f}s, snn::step() loop {sec:.2f}s on 1 of {os.cpu_count()} host cores "
                             f"(the reference is single-threaded)",
                   "sim_s_per_wall_s": args.cpu_baseline_steps * DT / sec}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref not present"}

    line = {
        "metric": METRIC, "value": events_total / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"brunel-c5-share: Brunel p={P_CONN}, in-degree-scaled weights, {n} neurons on {world} GPU(s) "
                               f"(BASELINE configs[4] weak-scaled: ~5e9 synapses per GPU)",
                   "neurons": n, "synapses": int(synapses_total), "dt": DT, "delay_steps": 15, "window_steps": 15,
                   "l2": "per-window delivery streams ~0.8 GB of CSR rows per GPU (> 126 MB L2)", "mode": "deterministic"},
        "sim_s_per_wall_s": args.steps * DT / (ms_max * 1e-3),
        "build_s": build_s,
        "gpu_launches": int(launches),
        "clocks": clk,
        "e2e": e2e,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
