"""N > 1 host path on CPU: world_size-2 gloo processes shard the frame by sample range and
sum-reduce their accumulation buffers; the result must equal the single-process render.  The
"renderer" here is the oracle (tests may use it); the sharding / reduce code is the product's."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scheme_raytrace_b200.host import sharding, scenes

W, H, SPP, SEED = 24, 16, 6, 7


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    S = O.OracleScene(scenes.cfg1_weekend(W, H))
    b, e = sharding.sample_range(rank, world, SPP)
    img, _ = S.render(W, H, e - b, max_depth=20, seed=SEED, spp_begin=b, nthreads=1)
    t = torch.from_numpy(img)
    sharding.reduce_accumulators(t, dist, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_ranges_partition():
    for world in (1, 2, 3, 4, 8):
        for spp in (0, 1, 7, 500, 4096):
            rs = [sharding.sample_range(r, world, spp, 3) for r in range(world)]
            assert rs[0][0] == 3 and rs[-1][1] == 3 + spp
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1


def test_two_rank_gloo_matches_single(tmp_path, orc):
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, 29731 + os.getpid() % 500, out), nprocs=2, join=True)
    got = np.load(out)
    S = orc.OracleScene(scenes.cfg1_weekend(W, H))
    ref, _ = S.render(W, H, SPP, max_depth=20, seed=SEED, nthreads=1)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
