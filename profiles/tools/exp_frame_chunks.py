#!/usr/bin/env python3
"""Experiment (GPU): the ATRAC3 LP2 pipeline fed in frame chunks through the streaming API, so that the per-chunk
band / spectrum buffers stay L2-resident (126 MB) instead of making a round trip through HBM.
Prints ms per 10^6 frames and per-kernel times for several chunk lengths."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import atracdenc_b200 as ab
import bench

S, C, step, FT = 1024, 2, 1024, 977
d_all = bench.gen_pcm_device(torch, S, FT, step, C, 0)
res = {}
for Fc in (977, 122, 32, 16, 8, 4):
    enc = ab.Encoder(3, C, device=0)
    n_calls = (FT + Fc - 1) // Fc
    # contiguous chunks [S][Fc*1024][C] cut out of the full batch (real data, so the data-dependent kernels do real work)
    chunks = [d_all[:, k * Fc * step:(k + 1) * Fc * step].contiguous() for k in range(n_calls)]
    d_out = torch.empty((S, Fc, enc.unit_bytes), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(enc.cuda_stream)
    def run():
        enc.reset()
        for ch in chunks:
            enc.encode_device(ch.data_ptr(), S, ch.shape[1] // step, d_out.data_ptr())
    for _ in range(2):
        run()
    enc.sync(); torch.cuda.synchronize()
    enc.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    enc.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kms, kcnt = enc.kernel_times(6)
    res[Fc] = {"ms_per_batch": ms, "calls": n_calls, "kernels_ms": {bench.KIND_NAMES[k]: kms[k] / reps for k in range(6) if kcnt[k]}}
    print(Fc, json.dumps(res[Fc]), flush=True)
    enc.close()
    del chunks
