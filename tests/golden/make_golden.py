#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the reference itself (oracle/_ref, built from /root/reference
by oracle/Makefile).  Run on the authoring box only; the fixtures are committed so the GPU box —
which has no /root/reference — can still check against reference output.

  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import atde_testlib as tl  # noqa: E402


def main():
    assert tl.ref_lib() is not None, "needs oracle/_ref (make -C oracle ref)"
    # config 1: ATRAC1, 1 s mono sine through the reference's own PCM pump
    pcm = tl.config1_sine()
    payload, sizes = tl.ref_encode(1, 1, pcm, total=44100)
    units = tl.pad_units(payload, sizes, 212)
    np.savez_compressed(HERE / "at1_config1_sine_mono.npz", units=units, sizes=sizes)
    print("config1:", units.shape, "payload bytes", int(sizes.sum()))
    # ATRAC1 stereo, noise + tone + bursts (exercises short windows, BFU trimming, boost)
    S, F, C = 2, 24, 2
    x = tl.synth_streams(S, F, 512, C, seed=0xA7AC)
    units, sizes, masks = [], [], []
    for s in range(S):
        p, z = tl.ref_encode(1, C, x[s])
        units.append(tl.pad_units(p, z, 212)[:F * C].reshape(F, C, 212))
        sizes.append(z[:F * C].reshape(F, C))
        masks.append(tl.ref_at1_stages(C, x[s])["masks"])
    np.savez_compressed(HERE / "at1_stereo_bursts.npz", pcm=x, units=np.stack(units), sizes=np.stack(sizes),
                        masks=np.stack(masks))
    print("stereo bursts:", np.stack(units).shape, "short-window frames:", int((np.stack(masks) != 0).sum()))
    # ATRAC3 LP2 (132 kbit/s) and LP4 (66 kbit/s joint stereo): tones (tonal components), level steps
    # and clicks (gain control), noise; frames straight from TAtrac3Encoder's lambda.  PCM stored as
    # the int16 it was quantised from.
    for name, kbit in (("at3_lp2_stereo.npz", 0), ("at3_lp4_js_stereo.npz", 64)):
        F, C = 24, 2
        xs = np.stack([tl.synth_rich(F, 1024, C, seed=0xA7AC + i, kind=k) for i, k in enumerate(("mix", "tones", "steps"))])
        frames = np.stack([tl.ref_at3_stages(C, x.reshape(-1), kbit)[2] for x in xs])
        pcm16 = np.rint(xs * 32768).astype(np.int16)
        assert np.array_equal(pcm16.astype(np.float32) / np.float32(32768), xs)
        np.savez_compressed(HERE / name, pcm=pcm16, frames=frames, kbit=kbit)
        print(name, frames.shape)
    # ATRAC3plus (default GHA settings), stereo and mono: tones (tone blocks, envelopes), level steps, noise;
    # frames straight from TAt3PEnc's lambda.  PCM stored as the int16 it was quantised from.
    for name, C in (("at3p_stereo.npz", 2), ("at3p_mono.npz", 1)):
        F = 10
        xs = np.stack([tl.synth_rich(F, 2048, C, seed=0xA7AC + 16 + i, kind=k) for i, k in enumerate(("mix", "tones", "steps"))])
        frames = np.stack([tl.ref_at3p_stages(C, x.reshape(-1))["frames"] for x in xs])
        pcm16 = np.rint(xs * 32768).astype(np.int16)
        assert np.array_equal(pcm16.astype(np.float32) / np.float32(32768), xs)
        np.savez_compressed(HERE / name, pcm=pcm16, frames=frames)
        print(name, frames.shape)


if __name__ == "__main__":
    main()
