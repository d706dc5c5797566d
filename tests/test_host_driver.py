"""
CPU: the host-side round driver (chbin_b200.run_iteration), query sharding (owned_slots) and the per-round label
exchange (TorchComm) with an oracle-backed engine -- single process and world_size 2 over gloo.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import chbin_b200
import oracle
from chbin_b200 import synth

from fake_engine import OracleEngine


def _case(hard):
    return synth.make_contig_features(260, 4, 1, 12, seed=9, concentration=60.0 if hard else 2000.0)


@pytest.mark.parametrize("hard,window", [(False, 0), (True, 0), (True, 37)])
def test_round_driver_equals_sequential(hard, window):
    X, bins, _ = _case(hard)
    perms = oracle.draw_permutations(bins, 4, seed=0)
    ref, info = oracle.fit_cluster(X, 4, bins, None, 5, 4, perms=perms, return_info=True)
    eng = OracleEngine(X, bins, 4, 5, window=window)
    changed = []
    for it in range(4):
        nch, rounds = chbin_b200.run_iteration(eng, perms[it])
        changed.append(nch)
        assert rounds >= 1
        if nch == 0:
            break
    assert np.array_equal(eng.get_labels(), ref)
    assert changed == list(info["changed"])


def test_owned_slots_partition():
    for U in (0, 1, 7, 100, 17501):
        for world in (1, 2, 3, 8):
            spans = [chbin_b200.owned_slots(U, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == U
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, hard, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, bins, _ = _case(hard)
        perms = oracle.draw_permutations(bins, 3, seed=0)
        U = perms.shape[1]
        u0, u1 = chbin_b200.owned_slots(U, rank, world)
        eng = OracleEngine(X, bins, 4, 5, u0, u1, window=64)
        comm = chbin_b200.TorchComm()
        for it in range(3):
            nch, _ = chbin_b200.run_iteration(eng, perms[it], comm)
            if nch == 0:
                break
        q.put((rank, eng.get_labels(), eng.qps))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("hard", [False, True])
def test_sharded_world2_gloo(hard):
    X, bins, _ = _case(hard)
    perms = oracle.draw_permutations(bins, 3, seed=0)
    ref = oracle.fit_cluster(X, 4, bins, None, 5, 3, perms=perms)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, hard, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, labels, qps in res:
        assert np.array_equal(labels, ref), f"rank {rank} diverged from the sequential reference"
    assert all(r[2] > 0 for r in res), "both ranks must have solved QPs (work is sharded)"


def test_small_stages_are_not_sharded():
    """clustering.SHARD_MIN_PAIRS: 20k contigs x 50 bins runs whole on every rank, the BASELINE sizes that are asked to scale shard."""
    assert not chbin_b200.clustering.sharding_pays(17_500, 50, 8)       # config #2
    assert chbin_b200.clustering.sharding_pays(95_000, 100, 8)          # config #3
    assert chbin_b200.clustering.sharding_pays(950_000, 500, 2)         # config #4
    assert not chbin_b200.clustering.sharding_pays(950_000, 500, 1)     # one rank never shards
