"""CPU, world size 2, gloo: the N > 1 host logic of bench.py -- the image batch is sharded by image with no
data-path collective, every rank transforms only its own images, and the whole-job throughput is the sum
of the ranks' images over the slowest rank's time.  The per-image transform here is the C oracle (the CUDA
library needs a GPU); what is under test is the sharding and the reduction, not the arithmetic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rbepwt_b200 import shard


def test_shard_range_partitions():
    for total in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            ranges = [shard.shard_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle
        from rbepwt_b200 import synth

        lo, hi = shard.shard_range(total, world, rank)
        sums = []
        for i in range(lo, hi):  # image i is seeded by its GLOBAL index: sharding must not change results
            lab = synth.voronoi_labels(32, 32, 12, seed=500 + i)
            img = synth.piecewise_smooth_image(lab, seed=500 + i)
            out = c_oracle.run(img, lab, 8, "bior4.4", "easypath", True, ncoefs=64)
            sums.append([i, float(out["decoded"].sum()), float(out["psnr"])])
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array(sums))
        dist.barrier()
        # slowest rank defines the job time; images add up
        t = shard.max_over_ranks(1.0 + rank)
        thr = shard.whole_job_throughput(hi - lo, 3, 1.0 + rank)
        np.save(os.path.join(out_dir, "thr%d.npy" % rank), np.array([t, thr]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_cover_the_batch_and_reduce(tmp_path):
    total, world = 5, 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    rows = np.concatenate([np.load(tmp_path / ("rank%d.npy" % r)) for r in range(world)])
    assert sorted(rows[:, 0].astype(int)) == list(range(total))
    # same images, unsharded
    from oracle import c_oracle
    from rbepwt_b200 import synth

    for i, s, p in rows:
        lab = synth.voronoi_labels(32, 32, 12, seed=500 + int(i))
        img = synth.piecewise_smooth_image(lab, seed=500 + int(i))
        out = c_oracle.run(img, lab, 8, "bior4.4", "easypath", True, ncoefs=64)
        assert float(out["decoded"].sum()) == s and float(out["psnr"]) == p
    for r in range(world):
        t, thr = np.load(tmp_path / ("thr%d.npy" % r))
        assert t == 2.0 and thr == pytest.approx(total * 3 / 2.0)
