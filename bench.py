#!/usr/bin/env python3
"""
bench.py — throughput of the B200 encode hot path (frames/s) with roofline and CPU baseline.

  python bench.py --gpus N --steps K --warmup W            our arm
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU path

A "step" is one pass of the hot path over one batch of synthetic PCM (S streams x F frames,
44.1 kHz stereo, int16-quantised).  `value` = whole-job stereo frames/s with PCM resident in HBM
(device entry point, CUDA events on the library's stream); `e2e` = the same metric through
atde_encode_batch() with HOST (pinned) buffers, H2D/D2H inside the timed region.
Multi-GPU: one process per GPU (torchrun), streams sharded by rank, no data-path collective
(the path shards by stream) -> weak scaling; NCCL only for the barrier and the
max-over-ranks of the device time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

WORKLOADS = {
    # name: (codec id, frame samples, S, F, units/frame, unit bytes, algorithmic QMF+MDCT bytes per stereo frame)
    "atrac1_stereo_1e6": dict(codec=1, step=512, S=1024, F=977, alg_bytes=8192, kbit=0, fp32_ops=2 * 48600,
                              kernel="at1_analysis_kernel (QMF+transient+MDCT)",
                              settings="reference defaults (EWM_AUTO, bfuidxconst 0)",
                              desc="ATRAC1 encode, 10^6-frame synthetic stereo batch (BASELINE.json configs[1])"),
    "atrac3_lp2_stereo_1e6": dict(codec=3, step=1024, S=1024, F=977, alg_bytes=16384, kbit=0, fp32_ops=2 * 125000,
                                  kernel="at3_qmf_kernel + at3_mdct_kernel (QMF tree, gain modulation, MDCT-512 x4)",
                                  settings="reference defaults: LP2 132300 bit/s, gain control + tonal components on",
                                  desc="ATRAC3 LP2 (132 kbps) encode, 10^6-frame synthetic stereo batch (BASELINE.json configs[2])"),
    "atrac3_lp4_stereo_1p25e6": dict(codec=3, step=1024, S=1024, F=1221, alg_bytes=16384, kbit=64, fp32_ops=2 * 125000,
                                     kernel="at3_qmf_kernel + at3_mdct_kernel (QMF tree, M/S, gain modulation, MDCT-512 x4)",
                                     settings="LP4 66150 bit/s joint stereo, gain control + tonal components on",
                                     desc="ATRAC3 LP4 (66 kbps, joint-stereo) encode, 1.25*10^6 frames per GPU "
                                          "(BASELINE.json configs[3] is this shard on each of 8 GPUs)"),
    # BASELINE.json configs[4]: 10^6 frames (1024 streams x 977 frames; ~75 GB of HBM with the per-batch workspaces)
    "atrac3plus_stereo": dict(codec=4, step=2048, S=1024, F=977, alg_bytes=32768, kbit=0,
                              kernel="at3p_pqf_kernel + at3p_mdct_kernel (16-band PQF, MDCT-256 x16)",
                              settings="reference defaults: GHA_ENABLED (pass input, write tonal, write residual)",
                              desc="ATRAC3PLUS encode, 10^6-frame synthetic stereo batch (BASELINE.json configs[4])"),
}
DEFAULT_WORKLOAD = "atrac3_lp2_stereo_1e6"
KIND_NAMES = ["qmf_mdct", "loudness_scan", "alloc_quant_pack", "gain_envelope", "gain_curve", "tonal_scale"]
KIND_NAMES_AT3P = ["pqf_mdct", "-", "scale_quant_pack", "tone_search", "tone_filter", "-"]
METRIC = "ATRAC3 stereo frames/s at 1/2/4/8 B200; QMF+MDCT achieved HBM GB/s vs peak"


def fp32_view(wl, frames, kernel_ms, clocks):
    """The same kernel against the FP32 issue roofline (the un-fused QMF + MDCT arithmetic makes the pair
    FP32-issue bound, not HBM bound): un-fused fp32 operations per stereo frame (SURVEY.md's count: 48.6 k per ATRAC1
    channel-frame, 125 k per ATRAC3 channel-frame) / kernel time, against 148 SMs x 128 lanes x the SM clock."""
    ops = wl.get("fp32_ops")
    if not ops or not kernel_ms:
        return None
    mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = 148 * 128 * mhz * 1e6 / 1e12
    ach = ops * frames / (kernel_ms / 1000.0) / 1e12
    return {"achieved": ach, "peak": peak, "unit": "T un-fused fp32 op/s", "frac": ach / peak,
            "ops_per_stereo_frame": ops, "sm_mhz": mhz}


def read_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def gen_pcm_device(torch, S, F, step, C, rank):
    """Synthetic PCM generated on the device (noise + per-stream tone + periodic +20 dB bursts, one silent
    stream in 61), int16-quantised then /32768 like a 16-bit WAV through libsndfile."""
    g = torch.Generator(device="cuda")
    g.manual_seed(0xA7AC + rank)
    n = F * step
    x = torch.empty((S, n, C), dtype=torch.float32, device="cuda")
    t = torch.arange(n, device="cuda", dtype=torch.float32)
    blk = 64
    for s0 in range(0, S, blk):
        s1 = min(S, s0 + blk)
        sid = torch.arange(s0, s1, device="cuda", dtype=torch.float32)
        freq = 100.0 * (1 + (sid % 160))
        v = 0.025 * (2 * torch.rand((s1 - s0, n, C), generator=g, device="cuda") - 1)
        ph = 2 * torch.pi * freq[:, None] * t[None, :] / 44100.0
        for c in range(C):
            v[:, :, c] += 0.03 * torch.sin(ph + c)
        fr = v.view(s1 - s0, F, step, C)
        fr[:, ::7, :64, :] *= 10.0
        silent = ((sid.long() % 61) == 60)
        v[silent] = 0
        x[s0:s1] = torch.round(v * 32767).clamp_(-32768, 32767) / 32768.0
    return x


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own encoder (oracle/_ref) on the host cores
def _cpu_worker(args):
    codec, C, step, F, n_streams, seed, use_ref, kbit = args
    import numpy as np
    import atde_testlib as tl
    pcm = tl.synth_streams(n_streams, F, step, C, seed=seed)
    t0 = time.perf_counter()
    for s in range(n_streams):
        if use_ref:
            tl.ref_encode(codec, C, pcm[s].reshape(-1), bitrate_kbit=kbit)
        else:
            tl.port_at1_encode(C, pcm[s].reshape(-1))
    return time.perf_counter() - t0, n_streams * F


def cpu_reference_rate(wl, streams_per_core=8, frames=977):
    """Stereo frames/s of the reference CPU encoder using every host core (one process per core,
    whole streams each), PCM generation excluded from the timed region."""
    import multiprocessing as mp
    import atde_testlib as tl
    use_ref = tl.ref_lib() is not None
    cores = os.cpu_count() or 1
    if not use_ref and wl["codec"] != 1:
        raise SystemExit("bench.py: the ATRAC3 CPU arm needs oracle/_ref (the reference build); it did not travel")
    jobs = [(wl["codec"], 2, wl["step"], frames, streams_per_core, 1000 + i, use_ref, wl["kbit"]) for i in range(cores)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    return dict(value=total / busy, cores=cores, kind="reference" if use_ref else "port", frames=total, busy_s=busy,
                sample=f"{cores} processes x {streams_per_core} streams x {frames} stereo frames "
                       f"({total} frames, {busy:.1f} s busy, {wall:.1f} s wall incl. PCM generation)",
                unit="frames/s")


# ------------------------------------------------------------------------------------------------
# parity at benchmark scale: the bench's OWN device-generated PCM through the reference encoder
def _verify_worker(args):
    path, shape, idxs, codec, C, kbit, n_out, unit_bytes = args
    import numpy as np
    import atde_testlib as tl
    pcm = np.load(path, mmap_mode="r")
    res = []
    for i in idxs:
        x = np.ascontiguousarray(pcm[i]).reshape(-1)
        payload, sizes = tl.ref_encode(codec, C, x, bitrate_kbit=kbit)
        if codec == 1:                                   # WriteFrame payloads of varying length: the container pads to 212
            units = tl.pad_units(payload, sizes, unit_bytes)[:n_out]
        else:                                            # fixed-size frames; what follows n_out is the engine's drain
            units = payload[:n_out * unit_bytes].reshape(n_out, unit_bytes)
        res.append((i, np.ascontiguousarray(units)))
    return res


def verify_against_reference(wl, sample_pcm, sample_out, world=1):
    """Every unit of `sample_out` ([n][units][unit_bytes], this library's output for the streams `sample_pcm`
    [n][samples][C]) against the reference encoder (oracle/_ref: src/main.cpp:697-716's loop over the same PCM),
    one process per host core.  Returns the `parity` object of the bench line."""
    import multiprocessing as mp
    import tempfile
    import numpy as np
    import atde_testlib as tl
    if tl.ref_lib() is None:
        return {"checked": False, "why": "oracle/_ref/libatde_ref.so did not travel"}
    n, n_units, ub = sample_out.shape
    C = sample_pcm.shape[2]
    cores = max(1, (os.cpu_count() or 1) // max(1, world))
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory(prefix="atde_verify_") as td:
        path = os.path.join(td, "pcm.npy")
        np.save(path, sample_pcm)
        jobs = [(path, sample_pcm.shape, list(range(k, n, cores)), wl["codec"], C, wl["kbit"], n_units, ub)
                for k in range(min(cores, n))]
        with mp.get_context("spawn").Pool(len(jobs)) as pool:
            parts = pool.map(_verify_worker, jobs)
    bad_units, bad_streams, first = 0, 0, None
    for part in parts:
        for i, want in part:
            diff = (want != sample_out[i]).any(-1)
            k = int(diff.sum())
            if k:
                bad_units += k
                bad_streams += 1
                if first is None:
                    first = {"sample_stream": int(i), "unit": int(np.argmax(diff))}
    upf = 2 if wl["codec"] == 1 else 1
    return {"checked": True, "frames_checked": int(n * n_units // upf), "units_checked": int(n * n_units),
            "streams_checked": int(n), "mismatches": int(bad_units), "mismatching_streams": int(bad_streams),
            "first_mismatch": first, "oracle": "oracle/_ref (unmodified reference, src/main.cpp's PCM loop)",
            "host_processes": len(jobs), "seconds": round(time.perf_counter() - t0, 2)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    rates = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(wl, streams_per_core=4 if wl["codec"] == 1 else 2, frames=wl["F"])
        if i >= args.warmup:
            rates.append(r)
    value = sum(r["value"] for r in rates) / len(rates)
    cb = dict(rates[-1]); cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * sum(r["busy_s"] for r in rates) / len(rates),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "sample_per_step": cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def bind_host_cores(local, world):
    """Pins this rank's host threads (and therefore the first touch of its pinned staging buffers) to an equal share
    of the cores NVML reports as local to its GPU.  Returns a description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = [c for c in range(ncpu) if (words[c // 64] >> (c % 64)) & 1]
        # GPUs that share this affinity mask split it evenly
        n_gpu = pynvml.nvmlDeviceGetCount()
        peers = []
        for g in range(min(n_gpu, max(world, 1))):
            w2 = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(g), (ncpu + 63) // 64)
            if list(w2) == list(words):
                peers.append(g)
        allowed = sorted(os.sched_getaffinity(0))
        near = [c for c in near if c in allowed] or allowed
        k = peers.index(local) if local in peers else 0
        share = max(1, len(near) // max(1, len(peers)))
        mine = near[k * share:(k + 1) * share] or near
        os.sched_setaffinity(0, mine)
        return {"gpu_local_cores": f"{near[0]}-{near[-1]}", "gpus_sharing_them": len(peers), "bound_to": f"{mine[0]}-{mine[-1]}"}
    except Exception as ex:                                             # no NVML / not permitted: run unbound
        return {"bound_to": None, "why": str(ex)[:80]}


def measure_workload(name, args, env, *, steps, warmup, verify_stride, with_i16=True, verify_only=False):
    """One workload on this rank's GPU: device-resident loop (`value`), host-buffer loop through the C ABI (`e2e`, float
    and int16 ingest), parity of a stream sample against the reference.  Returns the pieces of the JSON line (rank 0)
    or None (other ranks)."""
    import torch
    import torch.distributed as dist
    import atracdenc_b200 as ab
    rank, world, local, barrier = env["rank"], env["world"], env["local"], env["barrier"]

    wl = WORKLOADS[name]
    S, F, step, C = args.streams or wl["S"], args.frames or wl["F"], wl["step"], 2
    enc = ab.Encoder(wl["codec"], C, bitrate=wl["kbit"] * 1024, device=local)
    units, ub = enc.units_per_frame, enc.unit_bytes
    d_pcm = gen_pcm_device(torch, S, F, step, C, rank)
    d_out = torch.empty((S, F, units, ub), dtype=torch.uint8, device="cuda")
    host_kind = "pinned"

    def host_empty(shape, dtype):
        # pinned like a real ingest buffer; if the box refuses that much locked memory (8 ranks x 8 GB), fall back to
        # pageable memory rather than lose the run — `e2e.host_buffers` says which
        nonlocal host_kind
        try:
            return torch.empty(shape, dtype=dtype, pin_memory=True)
        except RuntimeError:
            host_kind = "pageable"
            return torch.empty(shape, dtype=dtype)

    h_pcm = host_empty((S, F * step, C), torch.float32)
    h_pcm.copy_(d_pcm)
    h_out = host_empty((S, F, units, ub), torch.uint8)
    stream = torch.cuda.ExternalStream(enc.cuda_stream, device=torch.device("cuda", local))
    torch.cuda.synchronize()
    dev_ms = host_ms = host16_ms = 0.0
    kms, kcnt, launches, clocks = [0.0] * 6, [0] * 6, 0, None

    if not verify_only:
        # ---- device-resident loop: `value` ----
        for _ in range(warmup):
            enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
        enc.sync()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = enc.launch_count
        enc.set_profiling(True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
        e1.record(stream)
        enc.sync()
        barrier()
        dev_ms = e0.elapsed_time(e1)
        kms, kcnt = enc.kernel_times(6)
        enc.set_profiling(False)
        launches = enc.launch_count - launches0
        clocks = sampler.stop() if rank == 0 else None

        # ---- end-to-end loop through the host API: `e2e` ----
        enc.reset()
        for _ in range(max(2, warmup // 2)):      # (two: the first call of a stream and its continuation differ)
            enc.encode_ptr(h_pcm.data_ptr(), S, F, h_out.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            enc.encode_ptr(h_pcm.data_ptr(), S, F, h_out.data_ptr())
        torch.cuda.synchronize()
        host_ms = (time.perf_counter() - t0) * 1000.0
        barrier()
    # host and device entry points must agree byte for byte from the same (fresh) stream state
    enc.reset()
    fo = enc.output_frames(F)
    enc.encode_device(d_pcm.data_ptr(), S, F, d_out.data_ptr())
    enc.sync()
    dev_bytes = d_out.view(-1)[: S * fo * units * ub].cpu()
    enc.reset()
    enc.encode_ptr(h_pcm.data_ptr(), S, F, h_out.data_ptr())
    same = bool(torch.equal(h_out.view(-1)[: S * fo * units * ub], dev_bytes))

    # ---- parity at benchmark scale: a fixed 1-in-`stride` sample of THIS batch's streams through the reference ----
    parity = None
    if verify_stride > 0:
        import hashlib
        ids = list(range(0, S, verify_stride))
        idt = torch.tensor(ids, device="cuda")
        sample_pcm = d_pcm.index_select(0, idt).cpu().numpy()
        sample_out = dev_bytes.view(S, fo * units, ub).index_select(0, idt.cpu()).numpy()
        parity = verify_against_reference(wl, sample_pcm, sample_out, world)
        parity["sample"] = f"streams 0, {verify_stride}, 2*{verify_stride}, ... of the {S} streams of this rank's batch"
        parity["batch_output_sha256"] = hashlib.sha256(dev_bytes.numpy().tobytes()).hexdigest()[:16]
        del sample_pcm, sample_out
        if world > 1:
            pv = torch.tensor([parity.get("frames_checked", 0), parity.get("mismatches", 0) if parity["checked"] else -1],
                              dtype=torch.int64, device="cuda")
            allp = [torch.zeros_like(pv) for _ in range(world)]
            dist.all_gather(allp, pv)
            parity["per_rank"] = [{"frames_checked": int(a[0]), "mismatches": int(a[1])} for a in allp]
            parity["frames_checked"] = int(sum(int(a[0]) for a in allp))
            parity["mismatches"] = int(sum(int(a[1]) for a in allp))
    if verify_only:
        enc.close()
        return {"workload": wl["desc"], "n_gpus": world, "streams_per_gpu": S, "frames_per_stream": F,
                "host_and_device_outputs_equal": same, "parity": parity} if rank == 0 else None

    # ---- the same through the int16 ingest entry point: half the H2D bytes ----
    same16 = None
    if with_i16:
        del h_pcm                                                       # its pinned block is reused for the int16 copy
        h_pcm16 = host_empty((S, F * step, C), torch.int16)
        h_pcm16.copy_(torch.round(d_pcm * 32768.0).to(torch.int16))    # the synthetic PCM is int16-quantised: exact
        enc.reset()
        enc.encode_ptr_i16(h_pcm16.data_ptr(), S, F, h_out.data_ptr())
        same16 = bool(torch.equal(h_out.view(-1)[: S * fo * units * ub], dev_bytes))
        for _ in range(max(2, warmup // 2)):      # (two: the first call of a stream and its continuation differ)
            enc.encode_ptr_i16(h_pcm16.data_ptr(), S, F, h_out.data_ptr())
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            enc.encode_ptr_i16(h_pcm16.data_ptr(), S, F, h_out.data_ptr())
        torch.cuda.synchronize()
        host16_ms = (time.perf_counter() - t0) * 1000.0
        barrier()
        del h_pcm16

    t = torch.tensor([dev_ms, host_ms, host16_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, host_ms, host16_ms = float(t[0]), float(t[1]), float(t[2])

    # ---- the optional exchange step at the edges: NCCL scatter of PCM shards from rank 0, gather of
    #      the bitstream shards onto rank 0, timed on the device; not part of `value` (the path itself has no collective)
    collective = None
    if world > 1 and env.get("collectives") and name == DEFAULT_WORKLOAD:
        collective = measure_collectives(torch, dist, env, d_pcm, d_out[:, :fo].contiguous(), S)

    ceiling = h2d_ceiling(torch, dist, env) if name == DEFAULT_WORKLOAD else None
    res = None
    if rank == 0:
        frames_total = S * F * world
        value = frames_total * steps / (dev_ms / 1000.0)
        e2e = frames_total * steps / (host_ms / 1000.0)
        peak, peak_kind = read_peak()
        k1_ms = kms[0] / max(1, steps)               # per step: ATRAC3 launches two kernels of kind 0 (QMF, MDCT)
        alg_bytes = wl["alg_bytes"] * S * F
        achieved = alg_bytes / (k1_ms / 1000.0) / 1e9 if k1_ms > 0 else None
        traffic = None
        tp = ROOT / "profiles" / f"traffic_{name}.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        h2d = S * F * step * C * 4
        res = {
            "value": value, "ms_per_step": dev_ms / steps, "steps": steps, "warmup": warmup,
            "config": {"workload": wl["desc"], "streams_per_gpu": S, "frames_per_stream": F, "channels": C,
                       "parallelism": f"streams sharded over {world} GPU(s), no data-path collective",
                       "l2_policy": f"inputs larger than L2 ({S * F * step * C * 4 / 2**20:.0f} MiB PCM per step)",
                       "settings": wl["settings"],
                       "signal": "per stream: uniform noise 0.025 FS + sine 0.03 FS at 100*(1 + s mod 160) Hz, x10 (+20 dB) "
                                 "64-sample burst at the start of every 7th frame, every 61st stream silent, int16-quantised; "
                                 "amplitudes are 1/10 of SURVEY.md 8(d)'s proposal (0.25 / 0.3), which clips under the bursts"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": wl["kernel"], "peak_source": f"of {peak_kind}",
                         "kernel_ms": k1_ms, "alg_bytes_per_launch": alg_bytes,
                         "kernel_share_of_step": (kms[0] / dev_ms) if dev_ms else None,
                         "fp32": fp32_view(wl, S * F, k1_ms, clocks),
                         "kernels_ms_per_step": {(KIND_NAMES_AT3P if wl["codec"] == 4 else KIND_NAMES)[k]: kms[k] / max(1, steps)
                                                 for k in range(6) if kcnt[k]}},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": S * F * units * ub, "ms_per_step": host_ms / steps,
                    "host_and_device_outputs_equal": same, "host_buffers": host_kind,
                    "h2d_GBps_aggregate": world * h2d / (host_ms / steps / 1000.0) / 1e9,
                    "h2d_GBps_ceiling_all_ranks_copying": ceiling,
                    "host_binding": env.get("binding")},
            "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
        }
        if with_i16:
            res["e2e"]["i16"] = {"value": frames_total * steps / (host16_ms / 1000.0), "unit": "frames/s",
                                 "h2d_bytes_per_step": h2d // 2, "d2h_bytes_per_step": S * F * units * ub,
                                 "ms_per_step": host16_ms / steps, "outputs_equal_float_path": same16,
                                 "note": "atde_encode_batch_i16: int16 PCM in (what a WAV file holds), converted on the device"}
        if collective:
            res["collective_ms"] = collective
            res["value_with_collectives"] = frames_total / ((dev_ms / steps + collective["scatter_pcm_ms"] + collective["gather_units_ms"]) / 1000.0)
    enc.close()
    del d_pcm, d_out, h_out
    torch.cuda.empty_cache()
    return res


def h2d_ceiling(torch, dist, env):
    """Aggregate pinned host -> device copy rate with EVERY rank copying at once (1 GiB each, five times): the ceiling of
    `e2e` on this host — all ranks share its memory controllers and PCIe root complexes."""
    n = 1 << 30
    try:
        h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            d.copy_(h, non_blocking=True)
        env["barrier"]()
        t0 = time.perf_counter()
        for _ in range(5):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1000.0
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if env["world"] > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        del h, d
        return env["world"] * 5 * n / (float(t[0]) / 1000.0) / 1e9
    except RuntimeError:
        return None


def measure_collectives(torch, dist, env, d_pcm, d_units, S):
    """atracdenc_b200.sharding.scatter_pcm / gather_units over NCCL (NVLink): rank 0 holds the whole job's PCM on its
    GPU, every rank receives its stream shard, encodes (not timed here), and rank 0 collects the bitstream shards."""
    from atracdenc_b200 import sharding
    rank, world = env["rank"], env["world"]
    dev = d_pcm.device
    try:
        full = None
        if rank == 0:
            full = torch.empty((world * S,) + tuple(d_pcm.shape[1:]), dtype=d_pcm.dtype, device=dev)
            for r in range(world):
                full[r * S:(r + 1) * S].copy_(d_pcm)
        torch.cuda.synchronize(); dist.barrier()
        times = []
        for it in range(3):                                             # first pass warms NCCL's channels up
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            mine = sharding.scatter_pcm(full, world * S, tuple(d_pcm.shape[1:]), d_pcm.dtype, dev, src=0)
            e1.record()
            got = sharding.gather_units(d_units, world * S, dst=0)
            e2.record()
            torch.cuda.synchronize(); dist.barrier()
            times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        t = torch.tensor(times[1:], dtype=torch.float64, device=dev).mean(0)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = torch.tensor([1 if (rank != 0 or bool(torch.equal(mine, d_pcm))) else 0], device=dev)
        if rank == 0:
            ok[0] = int(bool(torch.equal(mine, d_pcm)) and tuple(got.shape) == (world * S,) + tuple(d_units.shape[1:])
                        and bool(torch.equal(got[:S], d_units)))
        else:
            full0 = None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sb = (world - 1) * d_pcm.numel() * d_pcm.element_size()
        gb = (world - 1) * d_units.numel()
        del full, mine, got
        torch.cuda.empty_cache()
        return {"scatter_pcm_ms": float(t[0]), "gather_units_ms": float(t[1]), "scatter_bytes": sb, "gather_bytes": gb,
                "scatter_GBps_rank0_egress": sb / (float(t[0]) / 1000.0) / 1e9, "gather_GBps_rank0_ingress": gb / max(1e-9, float(t[1]) / 1000.0) / 1e9,
                "payload_ok": bool(int(ok[0])), "backend": "nccl (torch.distributed send/recv)",
                "what": "rank 0 -> every rank: its PCM shard; every rank -> rank 0: its bitstream shard (atracdenc_b200/sharding.py)"}
    except RuntimeError as ex:                                          # e.g. not enough HBM on rank 0 for the whole job's PCM
        return {"error": str(ex)[:200]}


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    binding = bind_host_cores(local, world) if not args.no_bind else {"bound_to": None, "why": "--no-bind"}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env = {"rank": rank, "world": world, "local": local, "barrier": barrier, "binding": binding,
           "collectives": not args.no_collectives}
    if args.verify_only:
        r = measure_workload(args.workload, args, env, steps=0, warmup=0, verify_stride=args.verify_stride or 8, verify_only=True)
        if rank == 0:
            print(json.dumps(r), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    main_res = measure_workload(args.workload, args, env, steps=args.steps, warmup=args.warmup, verify_stride=args.verify_stride)
    others = {}
    if args.workload == DEFAULT_WORKLOAD and not args.no_other_workloads and not (args.streams or args.frames):
        # the other BASELINE.json configs, three timed steps each, so that the driver's record carries them too
        for name, stride in (("atrac1_stereo_1e6", 64), ("atrac3_lp4_stereo_1p25e6", 16), ("atrac3plus_stereo", 64)):
            try:
                r = measure_workload(name, args, env, steps=3, warmup=3, verify_stride=stride if args.verify_stride else 0)
            except Exception as ex:                                     # never lose the headline line to a side workload
                r = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"} if rank == 0 else None
            if rank == 0 and r is not None:
                if "error" not in r:
                    r = {"value": r["value"], "unit": "frames/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                         "workload": r["config"]["workload"], "settings": r["config"]["settings"],
                         "roofline": {k: r["roofline"][k] for k in ("achieved", "peak", "frac", "fp32", "kernel", "kernel_ms", "kernels_ms_per_step")},
                         "e2e": {k: v for k, v in r["e2e"].items() if k != "host_binding"}, "parity": r["parity"]}
                others[name] = r

    if rank == 0:
        line = {"metric": METRIC, "value": main_res["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": main_res["config"], "roofline": main_res["roofline"], "e2e": main_res["e2e"],
                "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"], "parity": main_res["parity"]}
        for k in ("collective_ms", "value_with_collectives"):
            if k in main_res:
                line[k] = main_res[k]
        if others:
            line["other_workloads"] = others
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate(WORKLOADS[args.workload], streams_per_core=8 if WORKLOADS[args.workload]["codec"] == 1 else 4)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU (debug)")
    ap.add_argument("--frames", type=int, default=0, help="override frames per stream (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the 3-step passes over the other BASELINE.json configs")
    ap.add_argument("--no-collectives", action="store_true", help="N > 1: skip the NCCL scatter / gather measurement")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the rank to its GPU's local cores")
    ap.add_argument("--verify-stride", type=int, default=8,
                    help="parity: every k-th stream of the batch goes through the reference encoder on the host cores "
                         "(default 8: >= 1.2*10^5 frames of the 10^6-frame batches; 1 = every frame; 0 = off)")
    ap.add_argument("--verify", dest="verify_only", action="store_true",
                    help="only generate the batch, encode it once and verify it against the reference (no timing)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
