"""The N > 1 path on CPU: two processes (torch.distributed, gloo backend, world_size 2), each owning the
targets of its rank as the B200 ranks do: contiguous target ranges per population, all incoming
synapses of the local targets, every rank drawing the whole Poisson stream and keeping its own neurons, spike ids
exchanged with an all-gather.  The ranks run the sharded oracle; rank 0 checks the gathered rasters, states and the
synaptic-event tally against the unsharded run, bit for bit.  (The B200 runtime batches the exchange per min-delay
window — tests/test_gpu_sim.py::test_two_ranks_one_device; the oracle exchanges per step, which is the same data.)"""
import os
import socket
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

KW = dict(N=1500, p=0.1, w_exc=np.float32(2.0 / 300), w_inh=np.float32(-10.0 / 300))
STEPS = 120


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, out):
    import torch.distributed as dist

    from oracle_lib import Oracle, brunel_oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = Oracle()
        net, pops = brunel_oracle(orc, rank=rank, world=world, **KW)
        raster = []  # [step][pop] -> global spike list
        for _ in range(STEPS):
            net.step_update()
            local = [net.local_spikes(p) for p in pops]
            gathered = [None] * world
            dist.all_gather_object(gathered, local)
            row = [np.concatenate([gathered[r][pi] for r in range(world)]) for pi in range(len(pops))]
            for pi, p in enumerate(pops):
                net.step_set_spikes(p, row[pi])
            net.step_deliver()
            raster.append(row)
        states = [net.neurons(p) for p in pops[1:]]
        all_states = [None] * world
        dist.all_gather_object(all_states, states)
        events = [None] * world
        dist.all_gather_object(events, net.events())
        if rank == 0:
            out.put((raster, all_states, sum(events)))
    finally:
        dist.destroy_process_group()


def test_two_process_gloo_spike_exchange(orc):
    import torch.multiprocessing as mp

    from oracle_lib import brunel_oracle

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    raster, all_states, events = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full, pops = brunel_oracle(orc, **KW)
    for s in range(STEPS):
        full.step()
        for pi, p in enumerate(pops):
            assert np.array_equal(raster[s][pi], full.spikes(p, 0)), (s, pi)
    for k, p in enumerate(pops[1:]):
        assert np.array_equal(np.concatenate([st[k] for st in all_states]), full.neurons(p))
    assert events == full.events()


# ---- the product's own host logic of the N > 1 path: static synapse-count load balancing --------------------------------
def _balance_rank_main(rank, world, port, out):
    """Every rank holds a share of an adj_list's edges; the in-degree histogram is all-reduced (what a launcher does before
    it builds the network) and cut into target ranges by the product's spice_balance_ranges (host arithmetic of
    libspice_b200.so, no device): every rank must arrive at the same bounds."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(HERE.parent))
    import spice2_b200 as sp

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)  # the same list on every rank, each takes its share
        n_dst = 5000
        dst = np.minimum((rng.random(300000) ** 2.5 * n_dst).astype(np.int64), n_dst - 1)
        mine = dst[rank::world]
        hist = torch.from_numpy(np.bincount(mine, minlength=n_dst).astype(np.int64))
        dist.all_reduce(hist)
        bounds = sp.balance_ranges(hist.numpy(), world)
        gathered = [None] * world
        dist.all_gather_object(gathered, bounds.tolist())
        if rank == 0:
            out.put((gathered, hist.numpy(), np.bincount(dst, minlength=n_dst)))
    finally:
        dist.destroy_process_group()


def test_two_process_gloo_in_degree_balance():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_balance_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, hist, full = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(hist, full)  # the all-reduced histogram is the whole list's
    assert gathered[0] == gathered[1]  # every rank cuts the same ranges
    b = gathered[0]
    assert b[0] == 0 and b[-1] == len(full) and 0 < b[1] < len(full) // 2  # the heavy targets come first: the cut is early
    load = [int((full[b[r]: b[r + 1]] + 1).sum()) for r in range(2)]
    assert abs(load[0] - load[1]) <= full.max() + 1
