"""world_size-2 gloo test of the N>1 host logic: streams sharded by rank, PCM scattered from rank 0,
each rank encodes its shard through the C ABI (emulated kernels on CPU), bitstreams gathered on
rank 0 and compared with a single-process encode."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))


def _worker(rank, world, port, S, F, C, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import atde_testlib as tl
    import atracdenc_b200 as ab
    from atracdenc_b200 import sharding
    lib = ab.load_library(tl.EMU_SO)
    full = torch.from_numpy(tl.synth_streams(S, F, 512, C, seed=21)) if rank == 0 else None
    mine = sharding.scatter_pcm(full, S, (F * 512, C), torch.float32, "cpu")
    lo, hi = sharding.shard_range(S, rank, world)
    assert mine.shape[0] == hi - lo
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=lib)
    out = torch.from_numpy(enc.encode(mine.numpy(), hi - lo))
    enc.close()
    t = sharding.max_over_ranks(float(rank + 1))
    assert t == float(world)
    allu = sharding.gather_units(out, S)
    if rank == 0:
        np.save(result_path, allu.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    from atracdenc_b200.sharding import shard_range
    for S in (1, 5, 8, 1024):
        for W in (1, 2, 3, 8):
            r = [shard_range(S, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == S
            assert all(r[i][1] == r[i + 1][0] for i in range(W - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_sharded_encode_matches_single(tmp_path):
    import atde_testlib as tl
    import atracdenc_b200 as ab
    tl.build_emu()
    S, F, C = 5, 6, 2
    res = tmp_path / "gathered.npy"
    mp.spawn(_worker, args=(2, 29517, S, F, C, str(res)), nprocs=2, join=True)
    lib = ab.load_library(tl.EMU_SO)
    enc = ab.Encoder(ab.CODEC_ATRAC1, C, lib=lib)
    want = enc.encode(tl.synth_streams(S, F, 512, C, seed=21), S)
    enc.close()
    assert np.array_equal(np.load(res), want)
