"""CPU oracle: NumPy restatement of `digifil` (DSPSR) + `splice` (SIGPROC) semantics.

TEST INFRASTRUCTURE ONLY.  Nothing under ``frb-baseband_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and there only as the checker / the timed CPU arm.

PARITY UNPINNED.  The arithmetic of the reference's baseband->filterbank path lives in two
external executables that are neither vendored under /root/reference nor installable
offline: ``digifil`` (DSPSR, fork pharaofranz/dspsr, advised commit b68528e15e8 with PSRCHIVE
79ed5c0821 -- /root/reference/INSTALL.md:37-39) and ``splice`` (SIGPROC, fork
pharaofranz/sigproc, unpinned -- /root/reference/README.md:8).  The reference holds no tests,
fixtures or golden vectors, so this oracle restates the published
algorithm of those tools and anchors on the reference's own call sites:

* digifil argv                  /root/reference/process_vdif.py:156-182
* leakage factor (freq_res)     /root/reference/process_vdif.py:162
* pol -> -d/-P mapping          /root/reference/process_vdif.py:163-176
* .hdr semantics (BW sign etc.) /root/reference/process_vdif.py:115-139
* frequency plan per IF         /root/reference/base2fil.sh:54,65,254,407-414
* splice order (highest first)  /root/reference/base2fil.sh:350,367,422
* input bit layout              /root/reference/spif2file.sh:31-98,181

It is pinned instead by analytic known-answer tests (tests/test_oracle_kat.py): exhaustive
2-bit LUT, tone -> predicted channel with LSB/USB mirror, Parseval, Gaussian -> 8-bit
mean 127.5 / sigma 21.25, splice == concatenate, header round trip.

All arithmetic is float64 unless ``dtype=np.float32`` is requested (used by the timed CPU
baseline, which mirrors digifil's single-precision FFTW path).
"""
from __future__ import annotations

import numpy as np

try:  # scipy's pocketfft is ~2x faster than numpy's for float32 and keeps float32
    import scipy.fft as _fft
except Exception:  # pragma: no cover
    _fft = np.fft

# ----------------------------------------------------------------------------------------
# VDIF format (VDIF 1.1 spec).  The reference only ever reads these
# fields through `vdif_print_headers` (/root/reference/base2fil.sh:130-147,
# /root/reference/extract_baseband_chunk.py:36-70).
# ----------------------------------------------------------------------------------------
VDIF_FILL_WORD = 0x11223344
OPT2BIT_LO = 1.0
OPT2BIT_HI = 3.3359          # standard VLBI optimal 4-level reconstruction ratio
LEVELS_2BIT = np.array([-OPT2BIT_HI, -OPT2BIT_LO, OPT2BIT_LO, OPT2BIT_HI])


def parse_vdif_header(words: np.ndarray) -> dict:
    """Decode the first 4 little-endian 32-bit words of a VDIF frame header."""
    w0, w1, w2, w3 = (int(w) for w in words[:4])
    return {
        "seconds": w0 & 0x3FFFFFFF,
        "legacy": (w0 >> 30) & 1,
        "invalid": (w0 >> 31) & 1,
        "frame_nr": w1 & 0xFFFFFF,
        "ref_epoch": (w1 >> 24) & 0x3F,
        "frame_bytes": (w2 & 0xFFFFFF) * 8,
        "log2_nchan": (w2 >> 24) & 0x1F,
        "version": (w2 >> 29) & 0x7,
        "station": w3 & 0xFFFF,
        "thread": (w3 >> 16) & 0x3FF,
        "nbit": ((w3 >> 26) & 0x1F) + 1,
        "complex": (w3 >> 31) & 1,
    }


def vdif_epoch_mjd(ref_epoch: int) -> int:
    """MJD of 00:00 UTC at the start of a VDIF reference epoch (half-years since 2000)."""
    year = 2000 + ref_epoch // 2
    month = 1 if ref_epoch % 2 == 0 else 7
    # Fliegel & Van Flandern
    a = (14 - month) // 12
    y = year + 4800 - a
    m = month + 12 * a - 3
    jdn = 1 + (153 * m + 2) // 5 + 365 * y + y // 4 - y // 100 + y // 400 - 32045
    return jdn - 2400001  # JDN at noon -> MJD at preceding midnight


JA98_WINDOW = 512            # samples per level-setting window of DSPSR's 2-bit unpacker


def ja98_levels(phi):
    """Jenet & Anderson (1998) dynamic 2-bit output levels for a window in which a fraction `phi` of the samples
    fell between the thresholds: with t = sqrt(2) erfinv(phi) (threshold in units of sigma),
    lo = E|x|  for |x| < t = sqrt(2/pi) (1 - exp(-t^2/2)) / phi,  hi = E|x| for |x| > t = sqrt(2/pi) exp(-t^2/2) / (1 - phi),
    in units of the window's sigma.  phi = 0.6667 (thresholds at 0.9674 sigma) gives lo 0.4473, hi 1.4991 (ratio 3.352,
    the static optimal ratio is 3.3359).  phi is clamped to [1/512, 1 - 1/512]."""
    from scipy.special import erfinv
    phi = np.clip(np.asarray(phi, np.float64), 1.0 / JA98_WINDOW, 1.0 - 1.0 / JA98_WINDOW)
    t = np.sqrt(2.0) * erfinv(phi)
    e = np.exp(-0.5 * t * t)
    return np.sqrt(2.0 / np.pi) * (1.0 - e) / phi, np.sqrt(2.0 / np.pi) * e / (1.0 - phi)


def decode_vdif(buf: np.ndarray, *, nbit: int = 2, header_bytes: int = 32,
                frame_bytes: int | None = None, mask_invalid: bool = True,
                offset8: float = 127.5, return_flags: bool = False, mode: str = "static"):
    """VDIF byte stream (2 channels, real) -> x[2, nsamp] float64.

    Follows digifil's VDIF reader + unpacker as restated in SURVEY.md Appendix A1/A2:
    positional frames, 32 B header stripped, ch0 -> pol 0, ch1 -> pol 1, 2-bit offset
    binary 0..3 -> (-hi,-lo,+lo,+hi) with the static optimal levels, 8-bit offset binary
    -> code-127.5.  Fault handling (Appendix D1 default): samples of frames with the invalid
    bit set are 0.0; every 32-bit payload word equal to the fill pattern 0x11223344 yields
    0.0 for the time samples it holds.
    """
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    if frame_bytes is None:
        frame_bytes = parse_vdif_header(buf[:16].view("<u4"))["frame_bytes"]
    nframes = buf.size // frame_bytes
    fr = buf[: nframes * frame_bytes].reshape(nframes, frame_bytes)
    w0 = fr[:, 0:4].copy().view("<u4")[:, 0]
    invalid = (w0 >> 31).astype(bool)
    payload = fr[:, header_bytes:]
    pbytes = payload.shape[1]
    words = np.ascontiguousarray(payload).view("<u4")            # [nframes, pbytes/4]
    fill = words == VDIF_FILL_WORD
    if nbit == 2:
        # byte: bits[1:0]=ch0 t, [3:2]=ch1 t, [5:4]=ch0 t+1, [7:6]=ch1 t+1
        p = payload
        c = np.empty((nframes, pbytes, 2, 2), dtype=np.uint8)    # [.., t-in-byte, ch]
        c[..., 0, 0] = p & 3
        c[..., 0, 1] = (p >> 2) & 3
        c[..., 1, 0] = (p >> 4) & 3
        c[..., 1, 1] = (p >> 6) & 3
        x = LEVELS_2BIT[c]                                        # float64
        x = x.reshape(nframes, pbytes * 2, 2)
        samp_per_word = 8
    elif nbit == 8:
        x = payload.astype(np.float64).reshape(nframes, pbytes // 2, 2) - offset8
        samp_per_word = 2
    elif nbit == 1:
        # what jive5ab writes for the 1-bit modes (/root/reference/spif2file.sh:58-61): bit 2t = ch0, bit 2t+1 = ch1 of time
        # sample t; 0 -> -1.0, 1 -> +1.0
        bits = np.unpackbits(np.ascontiguousarray(payload), axis=1, bitorder="little")     # [nframes, pbytes * 8]
        x = 2.0 * bits.reshape(nframes, pbytes * 4, 2).astype(np.float64) - 1.0
        samp_per_word = 16
    else:
        raise ValueError(f"unsupported VDIF nbit={nbit}")
    if mask_invalid:
        x[invalid] = 0.0
        fm = np.repeat(fill, samp_per_word, axis=1)               # [nframes, t]
        x[fm] = 0.0
    out = np.ascontiguousarray(x.reshape(-1, 2).T)
    if mode == "ja98":
        # SURVEY.md Appendix A2 / D2, the mode DSPSR's unpacker runs in behind `digifil -2` (excision off): per
        # polarisation and window of 512 consecutive samples the two output magnitudes follow the fraction of samples
        # between the thresholds.  Masked samples (0.0) are left out of the count and stay 0.0.
        assert nbit == 2
        n = out.shape[1] // JA98_WINDOW * JA98_WINDOW
        w = out[:, :n].reshape(2, -1, JA98_WINDOW)
        a = np.abs(w)
        valid = a > 0
        nlow = (valid & (a < 2.0)).sum(axis=2)
        nval = valid.sum(axis=2)
        lo, hi = ja98_levels(np.where(nval > 0, nlow / np.maximum(nval, 1), 0.5))
        mag = np.where(a < 2.0, lo[..., None], hi[..., None])
        out = out.copy()
        out[:, :n] = (np.sign(w) * np.where(valid, mag, 0.0)).reshape(2, n)
        # a tail shorter than one window keeps the static levels (never reached: blocks are multiples of 512)
    elif mode != "static":
        raise ValueError(mode)
    if return_flags:
        return out, {"invalid_frames": int(invalid.sum()), "fill_words": int(fill.sum())}
    return out


# ----------------------------------------------------------------------------------------
# digifil stages
# ----------------------------------------------------------------------------------------
def filterbank(x: np.ndarray, nchan: int, freq_res: int, dtype=np.float64) -> np.ndarray:
    """DSPSR convolving filterbank `-F nchan:freq_res` on one real polarisation.

    Per non-overlapping block of M = 2*nchan*freq_res samples: unnormalised forward real FFT,
    Nyquist bin discarded, the M/2 bins cut into nchan segments of freq_res bins, each
    segment sent through an unnormalised backward complex FFT (Appendix A3).  Returns
    y[t, chan] complex with t = block*freq_res + m.  Trailing samples that do not fill a
    block are dropped.
    """
    M = 2 * nchan * freq_res
    nblk = x.size // M
    cdt = np.complex64 if dtype == np.float32 else np.complex128
    xb = np.asarray(x[: nblk * M], dtype=dtype).reshape(nblk, M)
    X = _fft.rfft(xb, axis=1)[:, : M // 2]
    seg = X.reshape(nblk, nchan, freq_res)
    if freq_res > 1:
        y = _fft.ifft(seg, axis=2, norm="forward")               # unnormalised backward
    else:
        y = seg
    return np.ascontiguousarray(np.transpose(y, (0, 2, 1))).reshape(nblk * freq_res, nchan).astype(cdt, copy=False)


#: output products per pol mode.  'I' = -d1, 'PPQQ' = -d2, 'I2' = -d3, 'coherence' = -d4
#: (PP,QQ,Re PQ*,Im PQ*; /root/reference/process_vdif.py:169-171), 'P0'/'P1' = -P0/-P1,
#: 'IQUV' = north-star Stokes in a circular basis (I=RR+LL, Q=2Re RL*, U=2Im RL*, V=RR-LL).
POL_MODES = ("P0", "P1", "I", "I2", "coherence", "IQUV", "PPQQ")


def pol_mode_from_reference(pol: int) -> str:
    """--pol value of process_vdif (/root/reference/process_vdif.py:58-64,163-176)."""
    return {0: "P0", 1: "P1", 2: "I", 3: "I2", 4: "coherence"}[pol]


def detect(yP: np.ndarray, yQ: np.ndarray, mode: str) -> np.ndarray:
    """Detection (Appendix A5).  Returns d[t, npol_out, chan] real."""
    pp = yP.real ** 2 + yP.imag ** 2
    qq = yQ.real ** 2 + yQ.imag ** 2
    if mode == "P0":
        out = [pp]
    elif mode == "P1":
        out = [qq]
    elif mode == "I":
        out = [pp + qq]
    elif mode == "I2":
        out = [(pp + qq) ** 2]
    elif mode == "PPQQ":
        out = [pp, qq]
    elif mode in ("coherence", "IQUV"):
        x = yP * np.conj(yQ)
        if mode == "coherence":
            out = [pp, qq, x.real, x.imag]
        else:
            out = [pp + qq, 2 * x.real, 2 * x.imag, pp - qq]
    else:
        raise ValueError(mode)
    return np.stack(out, axis=1)


def tscrunch(d: np.ndarray, D: int) -> np.ndarray:
    """Sum D consecutive time samples (Appendix A6); incomplete tail dropped."""
    if D <= 1:
        return d
    n = d.shape[0] // D
    return d[: n * D].reshape(n, D, *d.shape[1:]).sum(axis=1)


def rescale_stats(d: np.ndarray, nsamp_interval: int):
    """mean / sigma per (pol, chan) over the first interval (Appendix A7, `-c`)."""
    n = min(nsamp_interval, d.shape[0]) if nsamp_interval > 0 else d.shape[0]
    seg = d[:n].astype(np.float64)
    mean = seg.mean(axis=0)
    var = (seg * seg).mean(axis=0) - mean * mean
    scale = np.where(var > 0, 1.0 / np.sqrt(np.where(var > 0, var, 1.0)), 1.0)
    return mean, scale


def digitise(y: np.ndarray, nbit: int, digi_sigma: float = 6.0) -> np.ndarray:
    """SigProcDigitizer (Appendix A8): half the output range spans `digi_sigma` standard deviations (D9)."""
    if nbit == -32:
        return y.astype(np.float32)
    if nbit == 8:
        return np.clip(np.floor(y * (127.5 / digi_sigma) + 127.5 + 0.5), 0, 255).astype(np.uint8)
    if nbit == 16:
        return np.clip(np.floor(y * (32768.0 / digi_sigma) + 32768.0 + 0.5), 0, 65535).astype(np.uint16)
    if nbit == 2:
        q = np.clip(np.floor(y + 1.5 + 0.5), 0, 3).astype(np.uint8)
        flat = q.reshape(q.shape[0], -1)
        assert flat.shape[1] % 4 == 0
        f4 = flat.reshape(flat.shape[0], -1, 4)
        return (f4[..., 0] | (f4[..., 1] << 2) | (f4[..., 2] << 4) | (f4[..., 3] << 6)).astype(np.uint8)
    raise ValueError(f"nbit={nbit}")


DISPERSION_CONSTANT = 1.0 / 2.41e-4        # s MHz^2 pc^-1 cm^3 (DSPSR's 2.41e-4; same constant family as
                                           # the 8.3 us smearing rule of /root/reference/submit_job.py:62)


def chirp(nchan: int, freq_res: int, freq_mhz: float, bw_mhz: float, dm: float) -> np.ndarray:
    """In-channel coherent dedispersion response H[chan, bin].

    Channel c has sky centre f_c = FREQ - BW/2 + (c+0.5)*BW/N (BW signed); bin j of its segment
    sits at sky offset f = (j/freq_res - 0.5)*BW/N from that centre.
        H = exp(-i * 2pi * D * DM * 1e6 * f^2 / (f_c^2 (f_c + f))),  conjugated for LSB.
    The sign is the physical one for a forward FFT with exp(-i w t): a pulse dispersed by the
    cold-plasma delay D*DM/f^2 is compressed to ~1 sample (tests/test_oracle_kat.py::
    test_chirp_compresses_dispersed_pulse).  SURVEY A4 recalls the opposite sign for DSPSR;
    with it the same pulse is smeared to twice its dispersed width, so it is not used.
    """
    N, L = nchan, freq_res
    chbw = bw_mhz / N                                     # signed
    fc = freq_mhz - bw_mhz / 2 + (np.arange(N) + 0.5) * chbw
    j = np.arange(L)
    f = (j / L - 0.5) * chbw                              # sky offset from channel centre
    phase = 2 * np.pi * DISPERSION_CONSTANT * 1e6 * dm * f[None, :] ** 2 / (fc[:, None] ** 2 * (fc[:, None] + f[None, :]))
    if bw_mhz < 0:
        phase = -phase
    return np.exp(-1j * phase)


def smearing_samples(freq_mhz: float, bw_mhz: float, nchan: int, dm: float) -> float:
    """Dispersion smearing across the lowest channel of a subband, in channel samples
    (t = 8.3 us * DM * dnu_MHz / nu_GHz^3, the rule the reference uses in submit_job.py:62)."""
    chbw = abs(bw_mhz) / nchan
    f_low = freq_mhz - abs(bw_mhz) / 2 + chbw / 2
    t_us = 8.3 * dm * chbw / (f_low / 1000.0) ** 3
    return t_us / (nchan / abs(bw_mhz))                   # channel sample = nchan/|BW| us


def filterbank_dedisp(x: np.ndarray, nchan: int, freq_res: int, H: np.ndarray, nfilt_pos: int, nfilt_neg: int,
                      dtype=np.float64) -> np.ndarray:
    """Convolving filterbank with an in-channel response, overlap-save (Appendix A4): blocks of
    M = 2*nchan*freq_res samples advance by (freq_res - nfilt_pos - nfilt_neg)*2*nchan samples; each
    channel's segment is multiplied by H[c] before the backward FFT and the first nfilt_pos / last
    nfilt_neg samples of every block are discarded.  Returns y[t, chan]."""
    N, L = nchan, freq_res
    M = 2 * N * L
    keep = L - nfilt_pos - nfilt_neg
    step = keep * 2 * N
    nblk = 0 if x.size < M else (x.size - M) // step + 1
    out = np.empty((nblk * keep, N), dtype=np.complex128)
    for b in range(nblk):
        X = _fft.rfft(np.asarray(x[b * step: b * step + M], dtype=dtype))[: M // 2].reshape(N, L)
        y = _fft.ifft(X * H, axis=1, norm="forward")      # unnormalised backward
        out[b * keep:(b + 1) * keep] = y[:, nfilt_pos: L - nfilt_neg].T
    return out


def digifil(vdif: np.ndarray, *, freq_mhz: float, bw_mhz: float, nchan: int,
            freq_res: int | None = None, tscrunch_factor: int = 1, pol_mode: str = "I",
            out_nbit: int = 8, in_nbit: int = 2, start_s: float = 0.0, nsec: float | None = None,
            rescale_interval_s: float = 10.0, keep_bandpass: bool = False,
            frame_bytes: int | None = None, header_bytes: int = 32,
            dtype=np.float64, return_float: bool = False, dm: float = 0.0, coherent: bool = False,
            nfilt: tuple[int, int] | None = None, x: np.ndarray | None = None,
            decode_mode: str = "static", offset8: float = 127.5, digi_sigma: float = 6.0,
            fft_normalised: bool = False, rescale_mode: str = "constant") -> dict:
    """One IF: VDIF bytes -> SIGPROC samples, as `digifil -cont -c -b<nbit> -S<start> -T<nsec>
    -2 -D 0.0 [-t D] -d<..> -F<nchan>:<freq_res> [-I0]` (/root/reference/process_vdif.py:157-182).

    bw_mhz is signed like the .hdr BW keyword: negative = LSB
    (/root/reference/process_vdif.py:117-118).  Returns dict with 'data' [t, npol, chan] in
    file order (descending frequency), 'float' (pre-digitiser, natural channel order) when
    requested, and the header quantities tsamp_s / fch1 / foff / nchans / nifs / tstart.
    """
    if freq_res is None:
        freq_res = 512 if nchan <= 128 else 2 * nchan     # /root/reference/process_vdif.py:162
    vdif = np.ascontiguousarray(vdif, dtype=np.uint8)
    h0 = parse_vdif_header(vdif[:16].view("<u4"))
    if frame_bytes is None:
        frame_bytes = h0["frame_bytes"]
    fs = 2.0 * abs(bw_mhz) * 1e6                            # real Nyquist sampling
    spf = (frame_bytes - header_bytes) * 8 // (in_nbit * 2)  # time samples per frame
    fps = fs / spf
    f0 = int(round(start_s * fps))
    nfr = vdif.size // frame_bytes - f0
    if nsec is not None:
        nfr = min(nfr, int(round(nsec * fps)))
    if x is None:
        x = decode_vdif(vdif[f0 * frame_bytes: (f0 + nfr) * frame_bytes], nbit=in_nbit,
                        header_bytes=header_bytes, frame_bytes=frame_bytes, mode=decode_mode, offset8=offset8)
    # (x given: samples already decoded, e.g. by corner_turn(); vdif is then only read for its header time)
    if coherent and dm > 0:                                  # digifil -D dm -F nchan:D (process_vdif.py:177-180)
        H = chirp(nchan, freq_res, freq_mhz, bw_mhz, dm)
        yP = filterbank_dedisp(x[0], nchan, freq_res, H, nfilt[0], nfilt[1], dtype)
        yQ = filterbank_dedisp(x[1], nchan, freq_res, H, nfilt[0], nfilt[1], dtype)
    else:
        yP = filterbank(x[0], nchan, freq_res, dtype)
        yQ = filterbank(x[1], nchan, freq_res, dtype)
    d = tscrunch(detect(yP, yQ, pol_mode), tscrunch_factor)
    if fft_normalised:
        # D4: the unnormalised forward (M) and backward (freq_res) transforms leave detected power M * freq_res times the
        # input's; "normalised" divides that out (squared for the squared-intensity product).  Cancels under rescaling.
        norm = 1.0 / (2.0 * nchan * freq_res * freq_res)
        d = d * (norm * norm if pol_mode == "I2" else norm)
    tsamp = tscrunch_factor * nchan / (abs(bw_mhz) * 1e6)
    nint = int(np.floor(rescale_interval_s / tsamp + 0.5))
    if keep_bandpass:
        y = d.astype(np.float64)
        mean = np.zeros(d.shape[1:]); scale = np.ones(d.shape[1:])
    elif rescale_mode == "running" and nint > 0:
        # D8 alternative: without -c every interval is scaled with its own statistics
        y = np.empty(d.shape, np.float64)
        for r0 in range(0, d.shape[0], nint):
            mean, scale = rescale_stats(d[r0:r0 + nint], nint)
            y[r0:r0 + nint] = (d[r0:r0 + nint] - mean) * scale
    else:
        mean, scale = rescale_stats(d, nint)
        y = (d - mean) * scale
    if bw_mhz > 0:                                           # USB: flip so that foff < 0
        yo = y[:, :, ::-1]
    else:
        yo = y
    data = digitise(yo, out_nbit, digi_sigma)
    h_first = parse_vdif_header(vdif[f0 * frame_bytes: f0 * frame_bytes + 16].view("<u4"))
    tstart = (vdif_epoch_mjd(h_first["ref_epoch"]) + (h_first["seconds"] + h_first["frame_nr"] / fps) / 86400.0)
    out = {
        "data": data, "tsamp_s": tsamp, "nchans": nchan, "nifs": d.shape[1],
        "nbits": 32 if out_nbit == -32 else out_nbit,
        "fch1": freq_mhz + abs(bw_mhz) / 2 - abs(bw_mhz) / (2 * nchan),
        "foff": -abs(bw_mhz) / nchan, "tstart": tstart, "mean": mean, "scale": scale,
    }
    if return_float:
        out["float"] = d
    return out


def corner_turn(raw: np.ndarray, word_bits: int, bits, *, frame_bytes: int = 8032, header_bytes: int = 32,
                mask_invalid: bool = True, mark5b: bool = False, swap_sign_mag: bool = False) -> np.ndarray:
    """jive5ab spif2file as arithmetic (/root/reference/spif2file.sh:31-98,178-186): raw multi-BBC VDIF ->
    x[if, pol, t] decoded samples.  bits[if] = the 4 source bits of that IF's (pol0 lsb, pol0 msb, pol1 lsb,
    pol1 msb).  Samples of invalid frames and of 32-bit payload words equal to the fill pattern are 0.0.
    mark5b: Mark5B disk frames (spif2file.sh:79-93,105-108; pass frame_bytes 10016, header_bytes 16): frames without the
    sync word or with the test-vector bit carry no data.  swap_sign_mag: the first step of those modes' recipes,
    done literally -- the two bits of every 2-bit field change places before `bits` (the recipe as written) is applied."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    nframes = raw.size // frame_bytes
    fr = raw[: nframes * frame_bytes].reshape(nframes, frame_bytes)
    hw = fr[:, 0:8].copy().view("<u4")
    if mark5b:
        invalid = (hw[:, 0] != 0xABADDEED) | ((hw[:, 1] >> 15) & 1).astype(bool)
    else:
        invalid = (hw[:, 0] >> 31).astype(bool)
    pay = np.ascontiguousarray(fr[:, header_bytes:])
    dt = {16: "<u2", 32: "<u4", 64: "<u8"}[word_bits]
    w = pay.view(dt).astype(np.uint64)                                   # [frame, sample]
    if swap_sign_mag:
        even = np.uint64(0x5555555555555555)
        w = ((w & even) << np.uint64(1)) | ((w >> np.uint64(1)) & even)
    fill32 = pay.view("<u4") == VDIF_FILL_WORD
    if word_bits == 32:
        bad = fill32
    elif word_bits == 16:
        bad = np.repeat(fill32, 2, axis=1)
    else:
        bad = fill32[:, 0::2] | fill32[:, 1::2]
    bad = bad | invalid[:, None]
    out = np.empty((len(bits), 2, w.size), dtype=np.float64)
    for i, b in enumerate(bits):
        if len(b) == 2:                                                  # 1-bit samples (spif2file.sh:58-61): 0 -> -1, 1 -> +1
            x0 = 2.0 * ((w >> np.uint64(b[0])) & np.uint64(1)).astype(np.float64) - 1.0
            x1 = 2.0 * ((w >> np.uint64(b[1])) & np.uint64(1)).astype(np.float64) - 1.0
        else:
            c0 = ((w >> np.uint64(b[0])) & np.uint64(1)) | (((w >> np.uint64(b[1])) & np.uint64(1)) << np.uint64(1))
            c1 = ((w >> np.uint64(b[2])) & np.uint64(1)) | (((w >> np.uint64(b[3])) & np.uint64(1)) << np.uint64(1))
            x0, x1 = LEVELS_2BIT[c0.astype(np.int64)], LEVELS_2BIT[c1.astype(np.int64)]
        if mask_invalid:
            x0 = np.where(bad, 0.0, x0)
            x1 = np.where(bad, 0.0, x1)
        out[i, 0], out[i, 1] = x0.reshape(-1), x1.reshape(-1)
    return out


def splice(parts: list[dict]) -> dict:
    """SIGPROC `splice` (Appendix A10): inputs highest-frequency first
    (/root/reference/base2fil.sh:350,367); per time sample concatenate each file's row;
    header = first file's with nchans summed; shortest input ends the output."""
    n = min(p["data"].shape[0] for p in parts)
    for p in parts[1:]:
        assert p["nbits"] == parts[0]["nbits"] and abs(p["tsamp_s"] - parts[0]["tsamp_s"]) < 1e-15
    rows = [p["data"][:n].reshape(n, -1) for p in parts]
    out = dict(parts[0])
    out["data"] = np.concatenate(rows, axis=1)
    out["nchans"] = sum(p["nchans"] for p in parts)
    return out


def if_plan(nif: int, freq_lsb0: float, bw: float):
    """Frequency plan of base2fil (/root/reference/base2fil.sh:54,65,254,407-414):
    IF i (1-based) centred at freqLSB_0 + (i-1)*bw; odd = LSB, even = USB.  Returns
    [(if_number, centre_mhz, signed_bw)] in splice order (highest frequency first)."""
    plan = []
    for i in range(1, nif + 1):
        usb = (i % 2 == 0)
        plan.append((i, freq_lsb0 + (i - 1) * bw, bw if usb else -bw))
    return plan[::-1]


def base2fil(vdifs: dict[int, np.ndarray], *, nif: int, freq_lsb0: float, bw: float, **kw) -> dict:
    """All IFs of a scan -> spliced filterbank (digifil per IF, then splice)."""
    parts = []
    for i, fc, sbw in if_plan(nif, freq_lsb0, bw):
        parts.append(digifil(vdifs[i], freq_mhz=fc, bw_mhz=sbw, **kw))
    return splice(parts)
