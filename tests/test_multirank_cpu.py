"""CPU, world size 2 (gloo): the N>1 path of bench.py -- block partition of the cells, per-rank
counter-based inputs, one sum-reduce of the per-slice totals -- gives the same totals as one process
scheduling every cell.  The scheduling itself is done by the oracle here (no GPU in this container);
on the GPU box the same shard/reduce code runs over NCCL (bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radiosaber_b200 import shard, workload

S, UPS, G, T, CELLS, SEED = 6, 3, 64, 5, 10, 4


def _setup():
    w = np.full(S, 1.0 / S)
    p = np.tile(np.array([0, 0, 1, 1], dtype=np.int32), (S, 1))
    u2s = np.repeat(np.arange(S), UPS).astype(np.int32)
    return w, p, u2s


def _run_cells(first, n):
    from oracle.pyoracle import OracleScheduler
    w, p, u2s = _setup()
    o = OracleScheduler(9, w, p, u2s, n)
    _, dts = workload.tti_clock(T)
    for t in range(T):
        cqi = workload.synth_cqi(SEED, first, n, t, 1, len(u2s), G)[0]
        r2 = workload.synth_rand2(SEED, first, n, t, 1, S)[0]
        o.step(cqi, r2, dt=float(dts[t]))
    st = o.get_state()
    return shard.stats_from_state(st["cum_bytes"], st["cum_rbs"], u2s, S)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = shard.shard_cells(CELLS, world, rank)
    stats = _run_cells(first, n)
    t = torch.from_numpy(stats.view(np.int64).copy())
    dist.barrier()
    shard.reduce_stats(t, dst=0)
    if rank == 0:
        q.put(t.numpy().view(np.uint64).copy())
    dist.destroy_process_group()


def test_shard_cells_partition():
    for total in (1, 7, 4096, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_cells(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (f0, n0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + n0 == f1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


@pytest.mark.timeout(300)
def test_two_ranks_reduce_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _run_cells(0, CELLS)
    assert np.array_equal(got, want)
    assert want[0].sum() > 0
