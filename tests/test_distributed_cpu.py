"""CPU tests (-m "not gpu") of the N>1 host logic: image sharding + the detection all-gather over gloo with
world_size 2 (even and uneven shards)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from computervision.pytorch_b200 import distributed as cvd


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            ranges = [cvd.shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == cvd.shard_sizes(n, world)
    assert cvd.shard_sizes(1024, 8) == [128] * 8          # BASELINE config 5
    with pytest.raises(ValueError):
        cvd.shard_range(4, 2, 2)


def _fake_rows(n_global, max_det=5, width=7):
    """Deterministic per-image detection rows so that any rank can rebuild the expected global result."""
    g = torch.Generator().manual_seed(7)
    rows = torch.rand((n_global, max_det, width), generator=g)
    counts = torch.randint(0, max_det + 1, (n_global,), generator=g, dtype=torch.int32)
    for i in range(n_global):
        rows[i, int(counts[i]):] = 0
    return rows, counts


def _worker(rank, world, port, n_global, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, counts = _fake_rows(n_global)
        my_rows, my_counts = cvd.shard_batch([rows, counts], world, rank)
        all_rows, all_counts = cvd.gather_detections(my_rows.clone(), my_counts.clone(), n_global)
        ok = torch.equal(all_rows, rows) and torch.equal(all_counts, counts)
        per_image = cvd.split_rows(all_rows, all_counts)
        ok = ok and all(p.shape[0] == int(c) for p, c in zip(per_image, counts))
        if n_global % world == 0:   # the packed single-collective form (equal shards)
            packed = torch.cat((my_rows.reshape(-1), my_counts.to(torch.float32)))
            g = cvd.gather_packed(packed)
            r2, c2 = cvd.unpack_detections(g, n_global // world, rows.shape[1], rows.shape[2])
            ok = ok and torch.equal(r2, rows) and torch.equal(c2, counts)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [8, 7])
def test_gather_detections_world2_gloo(n_global):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, True), (1, True)]


def test_gather_without_process_group_is_identity():
    rows, counts = _fake_rows(3)
    r, c = cvd.gather_detections(rows, counts)
    assert r is rows and c is counts
