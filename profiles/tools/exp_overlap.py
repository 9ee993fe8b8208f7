#!/usr/bin/env python3
"""Experiment (GPU): does a pinned H2D copy run at full speed while the ATRAC3 kernels run?  Times (a) the
device-resident encode alone, (b) an 8.2 GB H2D copy alone, (c) both at once on different streams."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import atracdenc_b200 as ab
import bench
S, C, step, F = 1024, 2, 1024, 977
d_pcm = bench.gen_pcm_device(torch, S, F, step, C, 0)
d_out = torch.empty((S, F, 384), dtype=torch.uint8, device="cuda")
h = torch.empty((S, F * step, C), dtype=torch.float32, pin_memory=True)
d2 = torch.empty_like(d_pcm)
enc = ab.Encoder(3, C, device=0)
cs = torch.cuda.Stream()
for _ in range(2):
    enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
enc.sync()
def timed(do_enc, do_copy):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    if do_copy:
        with torch.cuda.stream(cs):
            e[0].record(cs); d2.copy_(h, non_blocking=True); e[1].record(cs)
    if do_enc:
        enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
    enc.sync(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1000
    return wall, (e[0].elapsed_time(e[1]) if do_copy else None)
for name, a, b in (("encode alone", True, False), ("copy alone", False, True), ("both", True, True), ("both", True, True)):
    w, c = timed(a, b)
    print(json.dumps({"case": name, "wall_ms": round(w, 1), "copy_ms": None if c is None else round(c, 1)}), flush=True)
