#!/usr/bin/env python3
"""Small scans through the input-format paths added late in round 2 (written for compute-sanitizer's memcheck; the tool is
closed on this GPU pool, so it runs as a plain smoke check of every new kernel variant on minimal inputs): 8-bit input on the
generic and dedispersion paths (stream-coordinate word masks, carried), 1-bit split and raw input, Mark5B frames (32- and
64-bit words, the latter with a gap placed by time code), dedispersion behind the generic kernels.  Faulty frames included;
prints a checksum per case.

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_inputs.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from frb_baseband_b200 import spif, synth  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402


def run(name, cfg, data, nframes):
    with Plan(cfg) as pl:
        cf, fb = int(pl.chunk_frames), cfg.frame_bytes
        out = []
        for f0 in range(0, nframes, cf):
            n = min(cf, nframes - f0)
            pl.push([data[f0 * fb:(f0 + n) * fb]])
            out.append(pl.pull().copy())
        pl.flush()
        out.append(pl.pull().copy())
        rows = pl.view_rows(np.concatenate([o for o in out if len(o)]))
        c = pl.counters()
    print(f"{name}: rows {rows.shape}, sum {float(np.asarray(rows, np.float64).sum()):.6e}, invalid {c['frames_invalid']}, "
          f"fill {c['frames_with_fill']}, missing {c['slots_missing']}", flush=True)


common = dict(out_nbit=-32, keep_bandpass=True)
faults = dict(invalid_frac=0.05, fill_frac=0.05)
# 8-bit: run-time generic kernels, compile-time generic kernels, L = 16 (register-resident columns), dedispersion
for nchan, L, nfr, chunk in ((64, 256, 130, 50), (8, 16, 60, 25), (512, 0, 560, 200)):
    v = synth.make_vdif(nfr, seed=3, bw_mhz=32.0, nbit=8, **faults)
    run(f"8-bit generic nchan {nchan} L {L}", PlanConfig(nchan=nchan, bw_mhz=[32.0], freq_res=L, tscrunch=4, in_nbit=8, chunk_units=chunk, **common), v, nfr)
v = synth.make_vdif(300, seed=4, bw_mhz=32.0, nbit=8, **faults)
run("8-bit dedispersion nchan 8", PlanConfig(nchan=8, bw_mhz=[-32.0], freq_mhz=[1400.0], tscrunch=4, in_nbit=8, dm=5.0, coherent=True, **common), v, 300)
# 1-bit split streams: tuned path, generic path; payload that is not a multiple of 32 bytes (scalar front end)
v = synth.make_vdif(40, seed=5, bw_mhz=32.0, nbit=1, **faults)
run("1-bit tuned nchan 32", PlanConfig(nchan=32, bw_mhz=[-32.0], tscrunch=8, in_nbit=1, **common), v, 40)
run("1-bit generic nchan 16 L 64", PlanConfig(nchan=16, bw_mhz=[-32.0], freq_res=64, tscrunch=2, in_nbit=1, chunk_units=15, **common), v, 40)
v = synth.make_vdif(64, seed=6, bw_mhz=32.0, nbit=1, payload_bytes=1000)
run("1-bit payload 1000", PlanConfig(nchan=8, bw_mhz=[-32.0], tscrunch=4, in_nbit=1, frame_bytes=1032, **common), v, 64)
# dedispersion behind the generic kernels: even and odd log2(freq_res), compile-time column kernel
v = synth.make_vdif(200, seed=7, bw_mhz=32.0, **faults)
for nchan, L, chunk in ((16, 64, 60), (16, 128, 60), (512, 0, 90)):
    run(f"generic dedispersion nchan {nchan} L {L}", PlanConfig(nchan=nchan, bw_mhz=[-32.0], freq_mhz=[1300.0], freq_res=L, tscrunch=2, dm=1.5,
                                                                  coherent=True, chunk_units=chunk, **common), v, 200)
# raw recordings: 1-bit VDIF, Mark5B with 32- and 64-bit words (the latter by time code, with a gap)
rng = np.random.default_rng(9)
for mode, nif, bw, by_header in (("VDIF_8000-1024-16-1", 8, 32.0, False), ("MARK5B-1024-16-2", 8, 16.0, False), ("MARK5B-2048-32-2", 16, 16.0, True)):
    W, bits = spif.recipe_for_mode(mode, nif)
    fb, hb, fmt = spif.frame_geometry(mode)
    spf = (fb - hb) * 8 // W
    nfr = 256
    nb = len(bits[0]) // 2
    codes = rng.integers(0, 1 << nb, size=(nif, 2, nfr * spf), dtype=np.uint8)
    if fmt:
        raw = synth.make_raw_mark5b(codes, W, bits, bw_mhz=bw, sec0=10).reshape(nfr, fb)
        raw[3, 0] ^= 1
        raw[4, 5] |= 0x80
    else:
        raw = synth.make_raw_vdif(codes, W, bits, bw_mhz=bw).reshape(nfr, fb)
        raw[3, 3] |= 0x80
    raw[6, hb + 64:hb + 128].view("<u4")[:] = 0x11223344
    if by_header:
        filler = np.repeat(raw[:1], 5, axis=0)
        filler[:, 8:12] = np.frombuffer(np.uint32(0x50000000).tobytes(), np.uint8)
        raw = np.concatenate([np.delete(raw, np.arange(100, 105), axis=0), filler])
    bws = [bw if i % 2 == 0 else -bw for i in range(1, nif + 1)]
    freqs = [1300.0 + (i - 1) * bw for i in range(1, nif + 1)]
    cfg = PlanConfig(nchan=8, bw_mhz=bws, freq_mhz=freqs, tscrunch=4, raw_word_bits=W, raw_bits=bits, raw_format=fmt, frame_bytes=fb,
                     header_bytes=hb, in_nbit=nb, frame_time_mode=int(by_header), **common)
    run(f"raw {mode}", cfg, raw.reshape(-1), nfr)
print("done")
