#!/usr/bin/env python3
"""Fills DESIGN.md's measurement table from bench JSON lines (usage: fill_design.py <n1.json> [<n8.json>])."""
import json, re, sys
def load(p):
    return json.loads(open(p).read().strip().splitlines()[-1])
def m(x):
    return f"{x / 1e6:.2f} M" if x >= 1e5 else f"{x / 1e3:.1f} k"
n1 = load(sys.argv[1])
n8 = load(sys.argv[2]) if len(sys.argv) > 2 else None
s = open("DESIGN.md").read()
def kern(k):
    return ", ".join(f"{a} {b:.1f}" for a, b in k.items())
rows = {"LP2": n1, "AT1": n1["other_workloads"]["atrac1_stereo_1e6"], "LP4": n1["other_workloads"]["atrac3_lp4_stereo_1p25e6"],
        "AT3P": n1["other_workloads"]["atrac3plus_stereo"]}
for tag, r in rows.items():
    s = s.replace(f"R2_{tag}_VALUE", m(r["value"])).replace(f"R2_{tag}_E2E", m(r["e2e"]["value"])).replace(f"R2_{tag}_I16", m(r["e2e"]["i16"]["value"]))
    s = s.replace(f"R2_{tag}_KERNELS", kern(r["roofline"]["kernels_ms_per_step"]) + f" (step {r['ms_per_step']:.0f} ms)")
s = s.replace("R2_LP2_REF", m(n1["cpu_baseline"]["value"]) + f" ({n1['e2e']['value'] / n1['cpu_baseline']['value']:.0f}x e2e)")
if n8:
    c = n8.get("collective_ms", {})
    o = n8["other_workloads"]
    txt = (f"Measured on the 8 x B200 box (one host socket, 32 cores, every GPU local to it; `bench.py` under torchrun, "
           f"each rank 1024 streams x 977 frames): **{m(n8['value'])} ATRAC3 LP2 frames/s device-resident** "
           f"({n8['value'] / 8 / n1['value'] * 100:.1f} % of 8 x the single-GPU value), end to end {m(n8['e2e']['value'])} from float PCM and "
           f"{m(n8['e2e']['i16']['value'])} from int16 PCM ({n8['e2e']['i16']['value'] / 8 / n1['e2e']['i16']['value'] * 100:.0f} % of 8 x single).  "
           f"The float path is bound by the HOST: the eight ranks pull {n8['e2e']['h2d_GBps_aggregate']:.0f} GB/s of PCM through one socket "
           f"(all ranks copying at once reach {n8['e2e'].get('h2d_GBps_ceiling_all_ranks_copying') or float('nan'):.0f} GB/s; one rank alone 55.6 GB/s), "
           f"where 8 x 36.7 = 294 GB/s would be needed to hide the copies behind the kernels; halving the bytes (int16 ingest) restores the scaling.  "
           f"configs[3] (LP4, 10^7 frames over 8 GPUs): {m(o['atrac3_lp4_stereo_1p25e6']['value'])} device-resident, "
           f"{m(o['atrac3_lp4_stereo_1p25e6']['e2e']['value'])} / {m(o['atrac3_lp4_stereo_1p25e6']['e2e']['i16']['value'])} end to end, "
           f"{o['atrac3_lp4_stereo_1p25e6']['parity']['frames_checked']} frames checked against the reference, {o['atrac3_lp4_stereo_1p25e6']['parity']['mismatches']} mismatches.  "
           f"ATRAC1 {m(o['atrac1_stereo_1e6']['value'])}, ATRAC3plus {m(o['atrac3plus_stereo']['value'])} frames/s on 8 GPUs.  "
           f"NCCL over NVLink: scattering the job's PCM from rank 0 ({c.get('scatter_bytes', 0) / 1e9:.1f} GB) takes {c.get('scatter_pcm_ms', 0):.1f} ms "
           f"({c.get('scatter_GBps_rank0_egress', 0):.0f} GB/s out of rank 0), gathering the bitstreams ({c.get('gather_bytes', 0) / 1e9:.2f} GB) "
           f"{c.get('gather_units_ms', 0):.1f} ms; with both inside the step the job runs at {m(n8.get('value_with_collectives', 0))} frames/s.")
    s = s.replace("R2_SCALING", txt)
open("DESIGN.md", "w").write(s)
