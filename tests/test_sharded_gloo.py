"""N > 1 host logic on CPU: two gloo ranks shard the rows, exchange packed top-K blocks with one
all_gather and merge them; the result must equal the single-rank answer.  The device kernels are
replaced by the CPU oracle here (this is a test of the sharding / packing / merge logic, which is
identical on NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
from oracle import network_oracle as orc
from smart_crossover.network_methods.sharded import block_views, row_partition


def test_row_partition_is_a_balanced_cover():
    for S, G in [(60000, 8), (20000, 3), (7, 8), (1, 1), (1001, 4)]:
        parts = [row_partition(S, G, g) for g in range(G)]
        assert parts[0][0] == 0 and sum(p[1] for p in parts) == S
        for a, b in zip(parts, parts[1:]):
            assert a[0] + a[1] == b[0]
        assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1


def test_balanced_row_bounds_cover_the_rows_in_proportion_to_the_rates():
    from smart_crossover.device import balanced_row_bounds
    rng = np.random.default_rng(0)
    for S, G in [(60000, 8), (20000, 8), (20000, 3), (301, 2), (10, 8), (8, 8), (100003, 5)]:
        spr = 1.0 + 0.05 * rng.random(G)                          # seconds per row of each GPU
        b = balanced_row_bounds(S, spr)
        assert b[0] == 0 and b[-1] == S and len(b) == G + 1
        assert all(y > x for x, y in zip(b, b[1:]))                 # every shard non-empty
        if S >= 1000 * G:
            rows = np.diff(b)
            t = rows * spr                                          # predicted time per GPU: equal within a tile row
            assert (t.max() - t.min()) / t.mean() < 2 * 16 * G / S + 1e-9
            assert all(v % 16 == 0 for v in b[1:-1])
    assert balanced_row_bounds(64, [1, 1]) == [0, 32, 64]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S, D, K, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, d, M = cases.ot_points(S, D, 123)
        y = cases.planted_duals(M, 123, 0.05)
        row0, S_loc = row_partition(S, world, rank)
        rc = orc.reduced_costs_ot(M[row0:row0 + S_loc], np.concatenate([y[row0:row0 + S_loc], y[S:]]))
        cnt, mn, ids, vals = orc.price_summary(rc, K)
        ids = ids + row0 * D                                    # global arc ids
        # the rank's block, laid out as device.Pricer.block: [K rc bits | K ids | header (4) | n_out, pad]
        block = torch.zeros(2 * K + 6, dtype=torch.int64)
        block[:K] = torch.full((K,), float("inf"), dtype=torch.float64).view(torch.int64)
        block[K:2 * K] = -1
        block[:ids.size] = torch.from_numpy(vals.copy()).view(torch.int64)
        block[K:K + ids.size] = torch.from_numpy(ids)
        block[2 * K] = cnt
        block[2 * K + 1] = int(np.float64(mn).view(np.int64))
        gathered = torch.empty(world, 2 * K + 6, dtype=torch.int64)
        dist.all_gather_into_tensor(gathered.view(-1), block)
        g_rc, g_id, hdr = block_views(gathered, K)
        count, cmax = hdr[:, 0].sum(), hdr[:, 0].max()
        g_rc, g_id = g_rc.contiguous(), g_id.contiguous()
        flat_rc, flat_id = g_rc.reshape(-1).numpy(), g_id.reshape(-1).numpy()
        real = flat_id >= 0
        o = np.lexsort((flat_id[real], flat_rc[real]))[:K]
        if rank == 0:
            ret["count"] = int(count)
            ret["cmax"] = int(cmax)
            ret["ids"] = flat_id[real][o]
            ret["rc"] = flat_rc[real][o]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [16, 300])
def test_two_rank_sharded_topk_equals_single_rank(K):
    S, D = 61, 47
    s, d, M = cases.ot_points(S, D, 123)
    y = cases.planted_duals(M, 123, 0.05)
    cnt, mn, ids, vals = orc.price_summary(orc.reduced_costs_ot(M, y), K)
    with mp.Manager() as manager:
        ret = manager.dict()
        mp.spawn(_worker, args=(2, _free_port(), S, D, K, ret), nprocs=2, join=True)
        assert ret["count"] == cnt and ret["cmax"] <= cnt
        assert np.array_equal(ret["ids"], ids) and np.array_equal(ret["rc"], vals)
