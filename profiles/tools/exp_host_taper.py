#!/usr/bin/env python3
"""Experiment (GPU): ATRAC3 host path with / without the tapered last chunks (ATDE_CHUNK_TAPER), float and int16 PCM."""
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
h_i16 = torch.empty((S, F * step, C), dtype=torch.int16, pin_memory=True); h_i16.copy_(torch.round(d_pcm * 32768.0).clamp_(-32768, 32767).to(torch.int16))
del d_pcm
h_out = torch.empty((S, F, 384), dtype=torch.uint8, pin_memory=True)
enc = ab.Encoder(3, C, device=0)
for mib in (sys.argv[1:] or ["768", "512"]):
    os.environ["ATDE_CHUNK_MIB"] = mib
    for taper in (1, 0, 1, 0):
        os.environ["ATDE_CHUNK_TAPER"] = str(taper)
        res = {}
        for kind, call, buf in (("f32", enc.encode_ptr, h_pcm), ("i16", enc.encode_ptr_i16, h_i16)):
            enc.reset()
            call(buf.data_ptr(), S, F, h_out.data_ptr())
            t0 = time.perf_counter()
            for _ in range(4): call(buf.data_ptr(), S, F, h_out.data_ptr())
            torch.cuda.synchronize(); res[kind] = round((time.perf_counter() - t0) * 1000 / 4, 1)
        print(json.dumps({"chunk_mib": mib, "taper": taper, "ms_per_batch": res}), flush=True)
enc.close()
