"""GPU parity across REAL devices: one process per GPU, spike rings mapped into the peers with CUDA IPC,
spike ids exchanged by NVLink peer stores (runtime.cu publish_window / wait_window), against the
single-process oracle.  Skipped with fewer than two GPUs (run with `gpurun --gpus 2`).

The in-process variant (tests/test_gpu_sim.py::test_two_ranks_one_device*) shares raw pointers on one
device; this one is the path bench.py takes under torchrun: cudaIpcGetMemHandle / cudaIpcOpenMemHandle,
stores that cross NVLink, flags read with system-scope fences.
"""
import multiprocessing as mp

import numpy as np
import pytest

from oracle_lib import brunel_oracle, flatten_raster, run_raster

pytestmark = pytest.mark.gpu

STEPS = 150


def _gpu_count():
    import ctypes

    try:
        cuda = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch

            return torch.cuda.device_count()
        except Exception:
            return 0
    n = ctypes.c_int(0)
    return n.value if cuda.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def _rank_main(rank, world, kw, steps, conns, inbox, outbox):
    """One rank = one process = one GPU.  Handles travel through the parent (any transport works)."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    try:
        from spice2_b200.samples import brunel

        net, pops = brunel(device=rank, rank=rank, world=world, **kw)
        net.finalize()
        outbox.put((rank, "handle", net.peer_handle()))
        net.set_peers(inbox.get())
        net.raster_enable(True)
        done = 0
        while done < steps:  # several run() calls: windows, readouts and the exchange interleave
            n = min(45, steps - done)
            net.step(n)
            done += n
        counts, ids = net.raster_read(steps)
        state = {pi: pops[pi].get_neurons().tobytes() for pi in (1, 2)}
        ranges = {pi: pops[pi].range() for pi in (1, 2)}
        ages = {age: [p.spikes(age) for p in pops] for age in (0, 14)}
        net.sync()
        outbox.put((rank, "result", dict(counts=counts, ids=ids, state=state, ranges=ranges, ages=ages,
                                         events=net.stats()["synaptic_events"])))
        inbox.get()  # stay alive (the peers' mappings of this rank's ring) until everyone is done
        net.close()
    except Exception as e:  # noqa: BLE001
        import traceback

        outbox.put((rank, "error", f"{e}\n{traceback.format_exc()}"))


def _run_ranks(world, kw, steps):
    ctx = mp.get_context("spawn")
    outbox = ctx.Queue()
    inboxes = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_rank_main, args=(r, world, kw, steps, None, inboxes[r], outbox)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        handles, results = {}, {}
        while len(handles) < world:
            r, kind, payload = outbox.get(timeout=600)
            assert kind == "handle", payload
            handles[r] = payload
        for q in inboxes:
            q.put([handles[r] for r in range(world)])
        while len(results) < world:
            r, kind, payload = outbox.get(timeout=900)
            assert kind == "result", payload
            results[r] = payload
        for q in inboxes:
            q.put("done")
        for p in procs:
            p.join(timeout=120)
        return [results[r] for r in range(world)]
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()


@pytest.mark.parametrize("plastic", [False, True])
def test_ranks_on_separate_devices_match_oracle(orc, plastic):
    world = min(_gpu_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    kw = dict(N=4000 * world // 2, p=0.1, w_exc=np.float32(2.0 / 400), w_inh=np.float32(-10.0 / 400), seed=(7,), plastic=plastic)
    onet, opops = brunel_oracle(orc, **kw)
    rows, ocounts = run_raster(onet, opops, STEPS)
    oids = flatten_raster(rows)
    res = _run_ranks(world, kw, STEPS)
    for r, out in enumerate(res):
        # every rank sees every spike: the whole raster on each of them
        assert np.array_equal(out["counts"], ocounts), r
        assert np.array_equal(out["ids"], oids), r
        for age in (0, 14):
            for p, op in enumerate(opops):
                assert np.array_equal(out["ages"][age][p], onet.spikes(op, age)), (r, age, p)
    for pi in (1, 2):
        got = b"".join(out["state"][pi] for out in res)
        want = onet.neurons(pi)
        if plastic:
            g = np.frombuffer(got, want.dtype)
            assert np.array_equal(g["Twait"], want["Twait"])
            assert np.allclose(g["V"], want["V"], rtol=0, atol=1e-5)
        else:
            assert got == want.tobytes()
    if not plastic:
        for _ in range(14):  # the reference tallies a spike's events delay - 1 steps after this backend does
            onet.step()
        assert sum(out["events"] for out in res) == onet.events()


def test_multi_rank_spec_across_two_devices():
    """tests/cpp/multi_rank_spec.cu with rank r on device r: per-synapse init hooks, a host-fed population and DeliverFromTo
    synapses (source snapshots stored into the peer over NVLink) reproduce the one-rank run."""
    import os
    import subprocess
    from pathlib import Path

    if _gpu_count() < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    root = Path(__file__).resolve().parent.parent
    exe = root / "tests" / "cpp" / "build" / "multi_rank_spec"
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(root / "tests" / "cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, timeout=600, env=dict(os.environ, SPICE_SPEC_TWO_DEVICES="1"))
    assert r.returncode == 0 and b"rank 1 on device 1" in r.stdout, r.stdout.decode() + r.stderr.decode()
