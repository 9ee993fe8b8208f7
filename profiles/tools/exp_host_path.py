#!/usr/bin/env python3
"""Experiment (GPU): where the host-buffer path (atde_encode_batch) loses time against the device-resident path.
For several chunk sizes: wall time per batch, summed kernel time per kind (CUDA events inside the library)."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import atracdenc_b200 as ab
import bench

S, C, step, F = 1024, 2, 1024, 977
d_pcm = bench.gen_pcm_device(torch, S, F, step, C, 0)
h_pcm = torch.empty((S, F * step, C), dtype=torch.float32, pin_memory=True); h_pcm.copy_(d_pcm)
h_out = torch.empty((S, F, 384), dtype=torch.uint8, pin_memory=True)
d_out = torch.empty((S, F, 384), dtype=torch.uint8, device="cuda")
for mib in (sys.argv[1:] or ["0", "192", "384", "768", "1536", "4096", "16384"]):
    if mib != "0":
        os.environ["ATDE_CHUNK_MIB"] = mib
    enc = ab.Encoder(3, C, device=0)
    if mib == "0":
        for _ in range(2): enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
        enc.sync(); enc.set_profiling(True); t0 = time.perf_counter()
        for _ in range(3): enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
        enc.sync(); ms = (time.perf_counter() - t0) * 1000 / 3
    else:
        for _ in range(2): enc.encode_ptr(h_pcm.data_ptr(), S, F, h_out.data_ptr())
        enc.set_profiling(True); t0 = time.perf_counter()
        for _ in range(3): enc.encode_ptr(h_pcm.data_ptr(), S, F, h_out.data_ptr())
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1000 / 3
    kms, kcnt = enc.kernel_times(6)
    print(json.dumps({"chunk_mib": mib if mib != "0" else "device-resident", "ms_per_batch": round(ms, 1),
                      "kernel_ms_sum": round(sum(kms) / 3, 1), "launches_timed": int(sum(kcnt) / 3),
                      "by_kind": {bench.KIND_NAMES[k]: round(kms[k] / 3, 1) for k in range(6)}}), flush=True)
    enc.close()
