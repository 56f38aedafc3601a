#!/usr/bin/env python
"""GPU time per simulation step of the reference's own sample networks (BASELINE configs[0], [2], [3]:
samples/brunel N = 20,000, samples/vogels N = 4,000, samples/brunel+ N = 20,000), next to the compiled
reference on one host core.  One JSON line per network -> profiles/samples_rNN.jsonl.

    python tools/bench_samples.py [--steps 3000] [--no-cpu]

These networks are launch-latency bound on a B200 (a step of brunel-20k is ~56 spikes x 1000 synapses);
the HBM-roofline discussion belongs to bench.py's scaled network.
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def gpu_run(build, steps, batch, mode=0):
    import torch

    net, pops = build(mode=mode) if mode else build()
    net.finalize()
    net.step(300)  # pre-roll: the E/I populations fire from step ~110 on
    net.sync()
    s0 = net.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream()
    net.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        e0.record(stream)
        done = 0
        while done < steps:
            n = min(batch, steps - done)
            net.step(n)
            done += n
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    s1 = net.stats()
    out = dict(ms_per_step=ms / steps, events_per_s=(s1["synaptic_events"] - s0["synaptic_events"]) / (ms * 1e-3),
               launches_per_step=(s1["kernel_launches"] - s0["kernel_launches"]) / steps)
    net.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import spice2_b200 as sp
    from spice2_b200.samples import brunel, vogels

    cases = [
        ("samples/brunel N=20000 (C1)", lambda **kw: brunel(**kw), dict(fn="brunel", kw={})),
        ("samples/vogels N=4000 (C3)", lambda **kw: vogels(**kw), dict(fn="vogels", kw={})),
        ("samples/brunel+ N=20000 (C4, plastic E->E)", lambda **kw: brunel(plastic=True, **kw), dict(fn="brunel", kw=dict(plastic=True))),
    ]
    for name, build, ref in cases:
        line = {"network": name, "steps": args.steps, "preroll": 300}
        line["gpu_run_call"] = gpu_run(build, args.steps, args.steps)          # one spice_run() for all steps
        line["gpu_step_by_step"] = gpu_run(build, min(args.steps, 1000), 1)    # snn::step() one at a time, no readout
        if "plastic" in name:
            line["gpu_fast_mode"] = gpu_run(build, args.steps, args.steps, mode=sp.MODE_FAST)
        if not args.no_cpu:
            from oracle_lib import RefShim

            if RefShim.available("fast"):
                shim = RefShim("fast")
                steps = args.steps if "plastic" not in name else min(args.steps, 300)
                t0 = time.time()
                r = getattr(shim, ref["fn"])(steps=steps, record=False, **ref["kw"])
                line["cpu_reference"] = dict(ms_per_step=r["sim_seconds"] * 1e3 / steps, build_s=r["build_seconds"], steps=steps,
                                             cores=1, note="compiled reference, its own flags, from step 0 (no pre-roll)")
                line["speedup_run_call"] = line["cpu_reference"]["ms_per_step"] / line["gpu_run_call"]["ms_per_step"]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
