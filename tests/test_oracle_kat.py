"""Analytic known-answer tests that pin the oracle (the reference ships no golden vectors;
SURVEY.md section 4 / 8c)."""
import numpy as np
import pytest

from oracle import digifil_oracle as o
from frb_baseband_b200 import synth, vdif
import algo_prototype as ap


def test_decode_lut_exhaustive():
    # every byte value, both nibbles: VDIF 2-bit offset binary -> (-hi,-lo,+lo,+hi)
    pay = np.resize(np.arange(256, dtype=np.uint8), 8000)
    fr = np.zeros(8032, np.uint8)
    fr[:32] = vdif.make_headers(1, frames_per_sec=4000).view(np.uint8)
    fr[32:] = pay
    x = o.decode_vdif(fr)
    lev = np.array([-3.3359, -1.0, 1.0, 3.3359])
    for k in range(8000):
        b = int(pay[k])
        assert x[0, 2 * k] == lev[b & 3] and x[1, 2 * k] == lev[(b >> 2) & 3]
        assert x[0, 2 * k + 1] == lev[(b >> 4) & 3] and x[1, 2 * k + 1] == lev[(b >> 6) & 3]


def test_decode_1bit_and_8bit_known_answers():
    """1-bit (spif2file.sh:58-61 mode): bit 2t = channel 0, bit 2t + 1 = channel 1 of time sample t, 0 -> -1, 1 -> +1;
    8-bit: offset binary, code - 127.5 (128 as the alternative)."""
    pay = np.resize(np.arange(256, dtype=np.uint8), 8000)
    fr = np.zeros(8032, np.uint8)
    fr[:32] = vdif.make_headers(1, frames_per_sec=2000, nbit=1).view(np.uint8)
    fr[32:] = pay
    x = o.decode_vdif(fr, nbit=1)
    assert x.shape == (2, 32000)
    for k in range(0, 8000, 37):
        b = int(pay[k])
        for t in range(4):
            assert x[0, 4 * k + t] == (1.0 if (b >> (2 * t)) & 1 else -1.0)
            assert x[1, 4 * k + t] == (1.0 if (b >> (2 * t + 1)) & 1 else -1.0)
    fr[:32] = vdif.make_headers(1, frames_per_sec=16000, nbit=8).view(np.uint8)
    x8 = o.decode_vdif(fr, nbit=8)
    assert x8.shape == (2, 4000) and x8[0, 0] == pay[0] - 127.5 and x8[1, 0] == pay[1] - 127.5 and x8[0, 3] == pay[6] - 127.5
    assert o.decode_vdif(fr, nbit=8, offset8=128.0)[1, 1] == pay[3] - 128.0
    # a fill word blanks the 16 time samples it holds (1-bit) / the 2 it holds (8-bit)
    fr[32 + 40:32 + 44] = np.frombuffer(np.uint32(vdif.FILL_WORD).tobytes(), np.uint8)
    assert np.all(o.decode_vdif(fr, nbit=8)[:, 20:22] == 0) and o.decode_vdif(fr, nbit=8)[0, 22] != 0
    fr[:32] = vdif.make_headers(1, frames_per_sec=2000, nbit=1).view(np.uint8)
    x1 = o.decode_vdif(fr, nbit=1)
    assert np.all(x1[:, 160:176] == 0) and np.all(np.abs(x1[:, 176:192]) == 1)


def test_decode_faults_zeroed():
    v = synth.make_vdif(64, seed=3, invalid_frac=0.2, fill_frac=0.2)
    x, fl = o.decode_vdif(v, return_flags=True)
    assert fl["invalid_frames"] > 0 and fl["fill_words"] > 0
    fr = v.reshape(64, 8032)
    inv = (fr[:, 0:4].copy().view("<u4")[:, 0] >> 31).astype(bool)
    xs = x.reshape(2, 64, 16000)
    assert np.all(xs[:, inv] == 0)
    words = np.ascontiguousarray(fr[:, 32:]).view("<u4")
    clean = ~inv & ~(words == vdif.FILL_WORD).any(axis=1)
    assert clean.any() and np.all(np.abs(xs[:, clean]) > 0)
    part = ~inv & (words == vdif.FILL_WORD).any(axis=1) & ~(words == vdif.FILL_WORD).all(axis=1)
    assert part.any()
    for f in np.nonzero(part)[0]:
        zero = np.repeat(words[f] == vdif.FILL_WORD, 8)
        assert np.all(xs[:, f, zero] == 0) and np.all(xs[:, f, ~zero] != 0)


@pytest.mark.parametrize("usb", [True, False])
def test_tone_lands_in_predicted_channel(usb):
    nchan, L = 32, 64
    M = 2 * nchan * L
    n = 4 * M
    c_true = 11
    # baseband frequency (c+0.5)/nchan of Nyquist
    t = np.arange(n)
    x = np.cos(np.pi * (c_true + 0.5) / nchan * t)
    y = o.filterbank(x, nchan, L)
    p = (np.abs(y) ** 2).sum(axis=0)
    assert p.argmax() == c_true
    assert p[c_true] > 100 * np.delete(p, c_true).max()
    # through digifil(): USB files are flipped so that the first channel is the highest frequency
    xs = np.stack([x, x]) * 2.0
    v = synth.make_vdif(n // 16000 + 1, seed=0, x=np.pad(xs, ((0, 0), (0, (n // 16000 + 1) * 16000 - n))))
    r = o.digifil(v, freq_mhz=1400.0, bw_mhz=32.0 if usb else -32.0, nchan=nchan, freq_res=L, out_nbit=-32,
                  keep_bandpass=True)
    prof = r["data"][:, 0, :].sum(axis=0)
    assert prof.argmax() == (nchan - 1 - c_true if usb else c_true)
    assert r["foff"] < 0 and abs(r["fch1"] - (1400.0 + 16.0 - 0.5)) < 1e-12


def test_parseval():
    rng = np.random.default_rng(1)
    nchan, L = 16, 32
    M = 2 * nchan * L
    x = rng.standard_normal(3 * M)
    y = o.filterbank(x, nchan, L)
    # unnormalised rFFT then unnormalised backward FFT of length L: sum|y|^2 = L * sum_k |X_k|^2
    # and sum over the kept half of the spectrum ~ M/2 * sum x^2 (DC/Nyquist edge terms aside)
    for b in range(3):
        X = np.fft.rfft(x[b * M:(b + 1) * M])[: M // 2]
        assert np.allclose((np.abs(y[b * L:(b + 1) * L]) ** 2).sum(), L * (np.abs(X) ** 2).sum(), rtol=1e-12)


def test_gaussian_requantise_moments():
    v = synth.make_vdif(256, seed=5, bw_mhz=16.0)
    r = o.digifil(v, freq_mhz=1400.0, bw_mhz=-16.0, nchan=32, tscrunch_factor=32, out_nbit=8)
    d = r["data"].astype(np.float64)
    assert abs(d.mean() - 127.5) < 0.6
    assert abs(d.std() - 127.5 / 6.0) < 0.6


def test_splice_is_concatenate_highest_first():
    nif = 4
    vd = {i: synth.make_vdif(256, seed=10 + i, bw_mhz=16.0) for i in range(1, nif + 1)}
    r = o.base2fil(vd, nif=nif, freq_lsb0=1300.0, bw=16.0, nchan=32, tscrunch_factor=32)
    assert r["nchans"] == nif * 32 and r["data"].shape[1] == nif * 32
    top = o.digifil(vd[4], freq_mhz=1300.0 + 3 * 16.0, bw_mhz=16.0, nchan=32, tscrunch_factor=32)
    assert np.array_equal(r["data"][:, :32], top["data"].reshape(-1, 32))
    assert abs(r["fch1"] - (1300.0 + 3.5 * 16.0 - 0.25)) < 1e-12


def test_if_plan_matches_reference_stepping():
    # base2fil.sh:54,65,254,407-414: odd IFs LSB from freqLSB_0 in steps of 2*bw, even IFs USB from +bw
    plan = o.if_plan(8, 1254.0, 32.0)
    assert [p[0] for p in plan] == [8, 7, 6, 5, 4, 3, 2, 1]
    assert plan[-1] == (1, 1254.0, -32.0) and plan[-2] == (2, 1286.0, 32.0)
    assert plan[0] == (8, 1254.0 + 7 * 32.0, 32.0)


@pytest.mark.parametrize("nchan,L", [(128, 512), (32, 512), (8, 16)])
def test_gpu_decomposition_equals_oracle(nchan, L):
    """The column-pass / row-pass / eps algebra the CUDA kernels implement is exactly the
    oracle's filterbank."""
    rng = np.random.default_rng(0)
    M = 2 * nchan * L
    x = rng.standard_normal((2, 2 * M + 7))
    yP, yQ = o.filterbank(x[0], nchan, L), o.filterbank(x[1], nchan, L)
    pP, pQ = ap.filterbank_dualpol(x, nchan, L)
    s = np.abs(yP).max()
    assert np.abs(pP - yP).max() < 1e-12 * s and np.abs(pQ - yQ).max() < 1e-12 * s


def test_epoch_mjd():
    assert o.vdif_epoch_mjd(0) == 51544            # 2000-01-01
    assert o.vdif_epoch_mjd(40) == 58849           # 2020-01-01
    assert vdif.epoch_mjd(41) == 59031             # 2020-07-01


@pytest.mark.parametrize("usb", [True, False])
def test_chirp_compresses_dispersed_pulse(usb):
    """Physical KAT for the dedispersion sign: an impulse dispersed by the cold-plasma law
    (delay = D*DM/f^2, DM 560) over a 32 MHz subband is compressed by the oracle's chirp to
    1-2 channel samples; with the opposite sign it stays smeared over hundreds."""
    BW, f_lo, DM, nchan, L = 32.0, 1238.0, 560.0, 128, 512
    M = 2 * nchan * L
    nblk = 6
    n = nblk * M
    k = np.arange(n // 2 + 1)
    fsky = f_lo + k * (2 * BW) / n
    phi = 2 * np.pi * o.DISPERSION_CONSTANT * DM * 1e6 * (1.0 / fsky - 1.0 / fsky[-1])
    x = np.fft.irfft(np.exp(1j * phi) * np.exp(-2j * np.pi * k * (n // 2) / n), n)
    if not usb:
        x = x * (-1.0) ** np.arange(n)                       # spectral inversion = the LSB-sampled view
    H = o.chirp(nchan, L, f_lo + BW / 2, BW if usb else -BW, DM)

    def width(c, Hc):
        best = None
        for b in range(nblk):
            Xb = np.fft.rfft(x[b * M:(b + 1) * M])[: M // 2].reshape(nchan, L)
            p = np.abs(np.fft.ifft(Xb[c] * Hc)) ** 2
            if best is None or p.max() > best.max():
                best = p
        return int((best > 0.5 * best.max()).sum())

    for c in (5, 64, 120):
        assert width(c, H[c]) <= 2
        assert width(c, np.conj(H[c])) > 50
    assert 100 < o.smearing_samples(f_lo + BW / 2, BW, nchan, DM) < 200      # ~152 samples at 1.238 GHz


def test_overlap_save_equals_long_convolution():
    """filterbank_dedisp keeps exactly the samples of each block that the response cannot wrap
    into: on the kept region it equals the same filter applied to a block shifted by half a step."""
    rng = np.random.default_rng(3)
    nchan, L = 8, 64
    M = 2 * nchan * L
    H = o.chirp(nchan, L, 1400.0, 16.0, 30.0)
    npos = nneg = 8
    x = rng.standard_normal(4 * M)
    y = o.filterbank_dedisp(x, nchan, L, H, npos, nneg)
    keep = L - npos - nneg
    assert y.shape == (((x.size - M) // (keep * 2 * nchan) + 1) * keep, nchan)
    # block 1 starts keep*2N samples in; its first kept sample follows block 0's last kept one
    y_shift = o.filterbank_dedisp(x[keep * 2 * nchan:], nchan, L, H, npos, nneg)
    assert np.allclose(y[keep:2 * keep], y_shift[:keep], rtol=0, atol=1e-9 * np.abs(y).max())


def test_filterbank_against_brute_force_dft():
    """Pins sign and normalisation conventions without any FFT library: X[k] = sum_n x[n] e^{-2 pi i k n/M}
    (unnormalised), channel c = bins [cL, (c+1)L), y_c[m] = sum_j X[cL+j] e^{+2 pi i j m / L} (unnormalised)."""
    rng = np.random.default_rng(9)
    nchan, L = 4, 8
    M = 2 * nchan * L
    x = rng.standard_normal(2 * M + 3)
    y = o.filterbank(x, nchan, L)
    n = np.arange(M)
    for b in range(2):
        xb = x[b * M:(b + 1) * M]
        X = np.array([(xb * np.exp(-2j * np.pi * k * n / M)).sum() for k in range(M // 2)])
        for c in range(nchan):
            for m in range(L):
                ref = sum(X[c * L + j] * np.exp(2j * np.pi * j * m / L) for j in range(L))
                assert abs(y[b * L + m, c] - ref) < 1e-9 * M * L
    assert y.shape == (2 * L, nchan)          # the 3 trailing samples are dropped


def test_digitiser_and_rescale_definitions():
    y = np.array([[-10.0, -6.0, -1.0, 0.0, 1.0, 5.999, 6.0, 10.0]])
    assert o.digitise(y, 8).tolist() == [[0, 0, 106, 128, 149, 255, 255, 255]]      # floor(y*21.25 + 128)
    assert o.digitise(y, 16).tolist() == [[0, 0, 27307, 32768, 38229, 65531, 65535, 65535]]
    assert o.digitise(np.array([[-3.0, -0.6, 0.4, 2.0]]), 2).tolist() == [[0b11_10_01_00]]
    d = np.arange(24, dtype=float).reshape(6, 1, 4)
    assert np.array_equal(o.tscrunch(d, 3)[:, 0, 0], [0 + 4 + 8, 12 + 16 + 20])      # sum, not mean
    mean, scale = o.rescale_stats(d, 4)                                               # first interval only
    assert np.allclose(mean[0], d[:4, 0].mean(axis=0)) and np.allclose(1 / scale[0], d[:4, 0].std(axis=0))


def _golden_scans():
    import base64
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_small_scans.json")))
    for c in g["cases"]:
        rows = np.frombuffer(base64.b64decode(c["rows_b64"]), np.uint8).reshape(c["shape"])
        yield c, rows, g["rescale_interval_s"]


def test_oracle_reproduces_committed_small_scans():
    """The seeded synthetic scans of tests/golden/oracle_small_scans.json (made by make_oracle_golden.py): same VDIF
    bytes from the seed, same 8-bit rows from the oracle (+-1 LSB allowed for a different FFT library build)."""
    import hashlib
    from frb_baseband_b200 import synth
    n = 0
    for c, rows, interval in _golden_scans():
        v = synth.make_vdif(c["nframes"], seed=c["seed"], bw_mhz=c["bw"], **c["sig"])
        assert hashlib.sha256(v.tobytes()).hexdigest() == c["vdif_sha256"], c["name"]
        sbw = c["bw"] if c["usb"] else -c["bw"]
        r = o.digifil(v, freq_mhz=1400.0, bw_mhz=sbw, nchan=c["nchan"], tscrunch_factor=c["D"], pol_mode=c["pol_mode"],
                      out_nbit=8, rescale_interval_s=interval)
        d = np.abs(r["data"].reshape(rows.shape).astype(int) - rows.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3, c["name"]
        assert r["fch1"] == pytest.approx(c["fch1"]) and r["tsamp_s"] == pytest.approx(c["tsamp_s"])
        n += 1
    assert n == 3


def test_ja98_levels_known_answer():
    """with the thresholds at 0.9674 sigma two thirds of the samples fall between them, and the
    conditional means of |x| are 0.4473 sigma (inner) and 1.4991 sigma (outer): ratio 3.352 against the static 3.3359."""
    lo, hi = o.ja98_levels(0.6667)
    assert abs(lo - 0.4473) < 2e-4 and abs(hi - 1.4991) < 2e-4
    v = synth.make_vdif(64, seed=12, bw_mhz=32.0)
    x = o.decode_vdif(v, mode="ja98")
    assert abs(x.std() - 0.94) < 0.02                       # quantisation keeps 88 % of the power: sigma 0.94
    xs = o.decode_vdif(v)
    assert np.array_equal(np.sign(x), np.sign(xs)) and np.array_equal(np.abs(x) > 1.0, np.abs(xs) > 2.0)


def test_knob_alternatives_change_only_what_they_name():
    v = synth.make_vdif(300, seed=13, bw_mhz=32.0, tone_frac=0.2)
    kw = dict(freq_mhz=1400.0, bw_mhz=-32.0, nchan=128, tscrunch_factor=16, rescale_interval_s=0.02)
    base = o.digifil(v, **kw)
    assert np.array_equal(o.digifil(v, fft_normalised=True, **kw)["data"], base["data"])       # cancels under -c
    wide = o.digifil(v, digi_sigma=3.0, **kw)["data"].astype(float)
    assert abs((wide - 127.5).std() / (base["data"].astype(float) - 127.5).std() - 2.0) < 0.1
    run = o.digifil(v, rescale_mode="running", **kw)["data"]
    nint = int(np.floor(0.02 / base["tsamp_s"] + 0.5))
    assert np.array_equal(run[:nint], base["data"][:nint]) and not np.array_equal(run[nint:], base["data"][nint:])
