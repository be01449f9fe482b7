#!/usr/bin/env python3
"""Stage-by-stage check of the round-2 channeliser (split and fused) against the NumPy model of the
decomposition (tests/algo_prototype.py) and the legacy kernels.  Prints, never asserts: one GPU run
shows where a discrepancy starts."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import algo_prototype as ap  # noqa: E402
from oracle import digifil_oracle as o  # noqa: E402
from frb_baseband_b200 import _lib, synth  # noqa: E402
from frb_baseband_b200.plan import Plan, PlanConfig  # noqa: E402


def run(path, v, nchan, bw, D, mode=_lib.POL_I, stages=False):
    os.environ["B2F_PATH"] = path
    cfg = PlanConfig(nchan=nchan, bw_mhz=[-bw] * len(v), tscrunch=D, pol_mode=mode, keep_bandpass=True, out_nbit=-32)
    out = {}
    with Plan(cfg) as pl:
        out["path"] = pl.path
        cf = int(pl.chunk_frames)
        fb = cfg.frame_bytes
        nfr = min(x.size // fb for x in v)
        rows = []
        for f0 in range(0, nfr, cf):
            n = min(cf, nfr - f0)
            pl.push([x[f0 * fb:(f0 + n) * fb] for x in v])
            if stages and f0 == 0:
                R, L = 2 * nchan, 512
                nblk = len(v) * (n * pl.geometry.samples_per_frame // (R * L))
                out["nblk"] = nblk
                out["colsum"] = pl.debug(5, np.complex64)[: nblk * R].reshape(nblk, R).copy()
                out["eps"] = pl.debug(6, np.complex64)[: nblk * nchan].reshape(nblk, nchan).copy()
                if pl.path == 1:
                    it = pl.debug(4, np.complex64)[: nblk * L * R].reshape(nblk, 32, R // 2, 16, 2)
                    out["inter"] = it.transpose(0, 1, 3, 2, 4).reshape(nblk, L, R).copy()
                elif pl.path == 0:
                    out["inter"] = pl.debug(4, np.complex64)[: nblk * L * R].reshape(nblk, L, R).copy()
                out["tstream"] = pl.debug(0, np.uint8).copy() if pl.path else None
            r = pl.pull()
            if len(r):
                rows.append(r.copy())
        pl.flush()
        r = pl.pull()
        if len(r):
            rows.append(r.copy())
        out["counters"] = pl.counters()
        out["rows"] = np.concatenate(rows).view(np.float32) if rows else np.empty((0, 0), np.float32)
    return out


def rel(a, b):
    a = np.asarray(a, np.complex128 if np.iscomplexobj(a) else np.float64)
    b = np.asarray(b, a.dtype)
    if a.shape != b.shape:
        return f"SHAPE {a.shape} vs {b.shape}"
    return f"{np.abs(a - b).max() / max(np.abs(b).max(), 1e-30):.2e}"


def main():
    cases = [(32, 16.0, 32, 1, 600), (128, 32.0, 16, 1, 1024), (8, 16.0, 4, 1, 300), (256, 32.0, 8, 1, 1024),
             (64, 32.0, 1, 1, 1024), (16, 16.0, 64, 1, 512), (128, 32.0, 512, 2, 2048)]
    if len(sys.argv) > 1:
        cases = cases[: int(sys.argv[1])]
    for nchan, bw, D, nif, nfr in cases:
        print(f"=== nchan {nchan} bw {bw} D {D} nif {nif} frames {nfr}", flush=True)
        v = [synth.make_vdif(nfr, seed=100 + i, bw_mhz=bw, tone_frac=0.3, invalid_frac=0.01 if i == 0 else 0, fill_frac=0.01 if i == 0 else 0)
             for i in range(nif)]
        R, L = 2 * nchan, 512
        M = R * L
        try:
            leg = run("legacy", v, nchan, bw, D, stages=True)
            spl = run("split", v, nchan, bw, D, stages=True)
            fus = run("fused", v, nchan, bw, D, stages=True)
        except Exception as e:          # noqa: BLE001
            print("  FAILED:", e, flush=True)
            continue
        print("  paths", leg["path"], spl["path"], fus["path"], "nblk", leg.get("nblk"))
        # transposed stream vs direct decode
        x = o.decode_vdif(v[0])
        nblk = spl["nblk"] // nif
        ts = spl["tstream"][: nblk * M].reshape(nblk, R // 2, 2, 32, 16)      # [blk][pair][half][lane][r]
        lut = np.zeros((17, 2), np.float32)
        lev = lambda k: (3.3359 if k in (0, 3) else 1.0) * (1 if k & 2 else -1)   # noqa: E731
        for c in range(16):
            lut[c] = (lev(c & 3), lev((c >> 2) & 3))
        dec = lut[ts >> 3]                                                    # [...][2 pols]
        lane = np.arange(32)
        item, col = lane >> 1, lane & 1
        bad = 0
        for b in (0, nblk - 1):
            zz = (x[0] + 1j * x[1])[b * M:(b + 1) * M].reshape(L, R)
            for half in (0, 1):
                for r in range(16):
                    rows_ = 32 * r + item + 16 * half
                    for pair in range(R // 2):
                        want = zz[rows_, 2 * pair + col]
                        got = dec[b, pair, half, :, r, 0] + 1j * dec[b, pair, half, :, r, 1]
                        bad += int(np.abs(want - got).max() > 0)
        print("  tstream mismatching groups:", bad)
        z = x[0] + 1j * x[1]
        for b in (0, nblk - 1):
            B, S = ap.column_pass(z[b * M:(b + 1) * M], R, L)
            e = ap.eps_from_colsum(S, R)
            print(f"  blk {b}: inter legacy {rel(leg['inter'][b], B)} split {rel(spl['inter'][b], B)} | colsum leg {rel(leg['colsum'][b], S)} "
                  f"split {rel(spl['colsum'][b], S)} fused {rel(fus['colsum'][b], S)} | eps leg {rel(leg['eps'][b], e)} split {rel(spl['eps'][b], e)} fused {rel(fus['eps'][b], e)}")
        print(f"  rows: split vs legacy {rel(spl['rows'], leg['rows'])}  fused vs legacy {rel(fus['rows'], leg['rows'])}  "
              f"fused == split bitwise: {np.array_equal(fus['rows'], spl['rows'])}  shape {fus['rows'].shape}")
        ck = ("frames_ok", "frames_invalid", "frames_with_fill", "fill_words", "frames_badhdr", "frames_misplaced")
        print("  counters legacy", [leg["counters"][k] for k in ck], "split", [spl["counters"][k] for k in ck], "fused", [fus["counters"][k] for k in ck], flush=True)


if __name__ == "__main__":
    main()
