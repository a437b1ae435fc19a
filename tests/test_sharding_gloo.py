"""CPU tier: the N>1 path (camera sharding + host-side gather of per-camera boxes) on gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcmtt_opticalflow_b200 import rig


def test_partition_is_exact():
    for n in (1, 4, 8, 16, 5):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                cams = rig.cameras_of_rank(n, world, r)
                assert all(rig.rank_of_camera(c, world) == r for c in cams)
                assert [rig.local_index(c, world) for c in cams] == list(range(len(cams)))
                seen += cams
            assert sorted(seen) == list(range(n))
    assert rig.cameras_of_rank(8, 8, 3) == [3] and rig.cameras_of_rank(16, 8, 3) == [3, 11]   # BASELINE C3 / C4


def _fake_boxes(cam):
    rng = np.random.default_rng(1000 + cam)
    k = int(rng.integers(1, 6))
    return np.arange(k, dtype=np.int32) + 10 * cam, rng.uniform(0, 500, size=(k, 4))


def _worker(rank, world, port, n_cams, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = {c: _fake_boxes(c) for c in rig.cameras_of_rank(n_cams, world, rank)}
        merged = rig.gather_results(dist, local, n_cams, dst=0)
        # the timed-region reduction bench.py uses: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ok = list(merged) == list(range(n_cams))
            for c in range(n_cams):
                ids, boxes = _fake_boxes(c)
                ok = ok and np.array_equal(merged[c][0], ids) and np.array_equal(merged[c][1], boxes)
            q.put((ok, float(t.item())))
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cams", [4, 5])
def test_gather_world_size_2(n_cams):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cams, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and tmax == 2.0


def test_single_process_gather_and_missing_camera():
    local = {c: _fake_boxes(c) for c in range(3)}
    merged = rig.gather_results(None, local, 3)
    assert list(merged) == [0, 1, 2]
    with pytest.raises(ValueError):
        rig.gather_results(None, local, 4)
