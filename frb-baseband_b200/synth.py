"""Seeded synthetic split-VDIF streams.

One stream per IF exactly as jive5ab's spif2file emits them
(/root/reference/spif2file.sh:178-186): frames of 32 B header + 8000 B payload, two
channels (ch0 = pol 0, ch1 = pol 1), 2 bit (or 8 bit) real samples, single thread,
non-legacy.  Samples: unit-variance Gaussian per pol quantised at {-0.9674, 0, +0.9674}
sigma to offset binary 0..3; optional tone, R/L-correlated component, dispersed impulse and
a fault mix (invalid-bit frames, fill-pattern payload words).
"""
from __future__ import annotations

import numpy as np

from . import vdif

THRESH_2BIT = 0.9674


def config_seed(config: int, if_index: int) -> int:
    return 20121102 + 1000 * config + if_index


def quantise_2bit(x: np.ndarray) -> np.ndarray:
    """float -> offset-binary code 0..3"""
    return ((x >= -THRESH_2BIT).astype(np.uint8) + (x >= 0).astype(np.uint8)
            + (x >= THRESH_2BIT).astype(np.uint8))


def pack_2bit(c0: np.ndarray, c1: np.ndarray) -> np.ndarray:
    """codes of ch0/ch1 (len n, n even) -> n/2 payload bytes (VDIF sample order)."""
    n = c0.size
    a0, a1 = c0.reshape(n // 2, 2), c1.reshape(n // 2, 2)
    return (a0[:, 0] | (a1[:, 0] << 2) | (a0[:, 1] << 4) | (a1[:, 1] << 6)).astype(np.uint8)


def make_signal(nsamp: int, rng: np.random.Generator, *, tone_frac: float | None = None,
                tone_amp: float = 0.5, rho: float = 0.0, impulse_at: int | None = None,
                impulse_amp: float = 30.0) -> np.ndarray:
    """x[2, nsamp] float64 before quantisation."""
    x = rng.standard_normal((2, nsamp))
    if rho:
        common = rng.standard_normal(nsamp)
        x = np.sqrt(1 - abs(rho)) * x + np.sqrt(abs(rho)) * common[None, :]
    if tone_frac is not None:
        t = np.arange(nsamp)
        # real sampled at fs: tone at tone_frac * (fs/2) baseband
        ph = np.pi * tone_frac * t
        x[0] += tone_amp * np.cos(ph)
        x[1] += tone_amp * np.sin(ph)
    if impulse_at is not None:
        x[:, impulse_at] += impulse_amp
    return x


def make_vdif(nframes: int, *, seed: int, bw_mhz: float = 32.0, nbit: int = 2,
              payload_bytes: int = 8000, sec0: int = 0, frame0: int = 0, ref_epoch: int = 40,
              invalid_frac: float = 0.0, fill_frac: float = 0.0, x: np.ndarray | None = None,
              **sig) -> np.ndarray:
    """Return a uint8 array holding `nframes` VDIF frames for one dual-pol IF."""
    rng = np.random.Generator(np.random.PCG64(seed))
    spf = payload_bytes * 8 // (nbit * 2)
    fps = int(round(2 * abs(bw_mhz) * 1e6 / spf))
    nsamp = nframes * spf
    if x is None:
        x = make_signal(nsamp, rng, **sig)
    if nbit == 2:
        payload = pack_2bit(quantise_2bit(x[0]), quantise_2bit(x[1]))
    elif nbit == 8:
        q = np.clip(np.floor(x * 20.0 + 128.0), 0, 255).astype(np.uint8)   # sigma = 20 counts
        payload = np.ascontiguousarray(q.T).reshape(-1)
    elif nbit == 1:                                                         # sign bits, ch0 / ch1 interleaved from bit 0 up
        payload = np.packbits(np.ascontiguousarray((x > 0).T).reshape(-1), bitorder="little")
    else:
        raise ValueError(nbit)
    payload = payload.reshape(nframes, payload_bytes)
    hdr = vdif.make_headers(nframes, frames_per_sec=fps, payload_bytes=payload_bytes, nbit=nbit,
                            ref_epoch=ref_epoch, sec0=sec0, frame0=frame0)
    if invalid_frac > 0:
        bad = rng.random(nframes) < invalid_frac
        hdr[bad, 0] |= np.uint32(1 << 31)
    out = np.empty((nframes, vdif.HEADER_BYTES + payload_bytes), dtype=np.uint8)
    out[:, : vdif.HEADER_BYTES] = hdr.view(np.uint8).reshape(nframes, vdif.HEADER_BYTES)
    out[:, vdif.HEADER_BYTES:] = payload
    if fill_frac > 0:
        # recorder-substituted data: whole payload (or a run of words) = fill pattern
        words = out[:, vdif.HEADER_BYTES:].view("<u4")
        bad = np.nonzero(rng.random(nframes) < fill_frac)[0]
        for k, f in enumerate(bad):
            if k % 2 == 0:
                words[f, :] = vdif.FILL_WORD
            else:                      # partial: a run of words inside the frame
                a = int(rng.integers(0, words.shape[1] - 64))
                words[f, a: a + int(rng.integers(1, 64))] = vdif.FILL_WORD
    return out.reshape(-1)


def if_file_name(exp: str, st: str, scan: str, i: int) -> str:
    """/root/reference/base2fil.sh:336,353"""
    return f"{exp}_{st}_no0{scan}_IF{i}.vdif"


def make_raw_mark5b(codes: np.ndarray, word_bits: int, bits, *, bw_mhz: float, mjd: int = 60000, sec0: int = 0,
                    frame0: int = 0) -> np.ndarray:
    """Raw multi-BBC Mark5B stream: like make_raw_vdif with 16-byte Mark5B headers and 10000-byte payloads.  `bits` are the
    source bits in the recorded word, i.e. what spif.parse_recipe returns for a swap_sign_mag recipe."""
    nif, _, nsamp = codes.shape
    dt = {16: np.uint16, 32: np.uint32, 64: np.uint64}[word_bits]
    w = np.zeros(nsamp, dtype=dt)
    for i in range(nif):
        nib = (codes[i, 0].astype(np.uint64) | (codes[i, 1].astype(np.uint64) << np.uint64(2)))
        for k in range(4):
            w |= (((nib >> np.uint64(k)) & np.uint64(1)) << np.uint64(bits[i][k])).astype(dt)
    pb, hb = vdif.MARK5B_PAYLOAD_BYTES, vdif.MARK5B_HEADER_BYTES
    spf = pb * 8 // word_bits
    nframes = nsamp // spf
    fps = int(round(2 * abs(bw_mhz) * 1e6 / spf))
    hdr = vdif.make_mark5b_headers(nframes, frames_per_sec=fps, mjd=mjd, sec0=sec0, frame0=frame0)
    out = np.empty((nframes, hb + pb), dtype=np.uint8)
    out[:, :hb] = hdr.view(np.uint8).reshape(nframes, hb)
    out[:, hb:] = w[: nframes * spf].view(np.uint8).reshape(nframes, pb)
    return out.reshape(-1)


def make_raw_vdif(codes: np.ndarray, word_bits: int, bits, *, bw_mhz: float, payload_bytes: int = 8000,
                  ref_epoch: int = 40, sec0: int = 0, frame0: int = 0) -> np.ndarray:
    """Raw multi-BBC VDIF stream (what the recorder holds): codes[if, pol, t] in 0..3 are scattered to
    the bit positions `bits[if] = [pol0 lsb, pol0 msb, pol1 lsb, pol1 msb]` of one word_bits-bit word per
    time sample -- the inverse of the spif2file recipe."""
    nif, _, nsamp = codes.shape
    dt = {16: np.uint16, 32: np.uint32, 64: np.uint64}[word_bits]
    w = np.zeros(nsamp, dtype=dt)
    sample_bits = len(bits[0]) // 2                 # two recipe entries per IF: 1-bit samples, codes in 0..1
    for i in range(nif):
        nib = (codes[i, 0].astype(np.uint64) | (codes[i, 1].astype(np.uint64) << np.uint64(sample_bits)))
        for k in range(2 * sample_bits):
            w |= (((nib >> np.uint64(k)) & np.uint64(1)) << np.uint64(bits[i][k])).astype(dt)
    spf = payload_bytes * 8 // word_bits
    nframes = nsamp // spf
    fps = int(round(2 * abs(bw_mhz) * 1e6 / spf))
    hdr = vdif.make_headers(nframes, frames_per_sec=fps, payload_bytes=payload_bytes, nbit=sample_bits,
                            log2_nchan=int(np.log2(max(1, word_bits // sample_bits))), ref_epoch=ref_epoch, sec0=sec0, frame0=frame0)
    out = np.empty((nframes, vdif.HEADER_BYTES + payload_bytes), dtype=np.uint8)
    out[:, : vdif.HEADER_BYTES] = hdr.view(np.uint8).reshape(nframes, vdif.HEADER_BYTES)
    out[:, vdif.HEADER_BYTES:] = w[: nframes * spf].view(np.uint8).reshape(nframes, payload_bytes)
    return out.reshape(-1)
