"""The knobs for what cannot be pinned without a running digifil: each
alternative against the oracle run in the same mode.  The defaults are what every other test exercises."""
import numpy as np
import pytest

from oracle import digifil_oracle as o
from frb_baseband_b200 import _lib, synth
from frb_baseband_b200.plan import Plan, PlanConfig
from helpers import REL_TOL, assert_rel

pytestmark = pytest.mark.gpu


def _run(cfg, v, pulls_until_empty=False):
    out = []
    with Plan(cfg) as pl:
        fb, cf = cfg.frame_bytes, int(pl.chunk_frames)
        nfr = v.size // fb
        for f0 in range(0, nfr, cf):
            pl.push([v[f0 * fb:(f0 + min(cf, nfr - f0)) * fb]])
            while True:
                r = pl.pull()
                if not len(r):
                    break
                out.append(r.copy())
                if not pulls_until_empty:
                    break
        pl.flush()
        while True:
            r = pl.pull()
            if not len(r):
                break
            out.append(r.copy())
        path = pl.path
        rows = pl.view_rows(np.concatenate(out))
    return rows, path


@pytest.mark.parametrize("path", ["split", "fused"])
@pytest.mark.parametrize("nchan,bw,D", [(128, 32.0, 16), (32, 16.0, 32), (256, 32.0, 4)])
def test_ja98_dynamic_levels(gpu, monkeypatch, path, nchan, bw, D):
    """decode_mode = JA98 (what DSPSR's unpacker does behind `digifil -2`, process_vdif.py:157): float spectra within
    1e-5 of the oracle decoding with per-window levels; faulty frames stay out of the level statistics."""
    monkeypatch.setenv("B2F_PATH", path)
    v = synth.make_vdif(1024 if bw > 16 else 512, seed=4242, bw_mhz=bw, tone_frac=0.31, invalid_frac=0.01, fill_frac=0.01)
    cfg = PlanConfig(nchan=nchan, bw_mhz=[-bw], tscrunch=D, out_nbit=-32, keep_bandpass=True, decode_mode=_lib.DECODE_JA98)
    rows, used = _run(cfg, v)
    assert used == {"split": 1, "fused": 2}[path]
    ref = o.digifil(v, freq_mhz=1400.0, bw_mhz=-bw, nchan=nchan, tscrunch_factor=D, out_nbit=-32, keep_bandpass=True,
                    decode_mode="ja98")["data"].astype(np.float64)
    assert_rel(rows.reshape(ref.shape), ref, REL_TOL, f"JA98 {path}")
    static = o.digifil(v, freq_mhz=1400.0, bw_mhz=-bw, nchan=nchan, tscrunch_factor=D, out_nbit=-32, keep_bandpass=True)["data"]
    assert np.abs(static / ref - 1).max() > 0.5              # and it is a different result from the static levels


@pytest.mark.parametrize("nchan,freq_res,D,nframes,chunk,dm", [(512, 0, 4, 300, 100, 0.0), (64, 256, 8, 120, 50, 0.0), (16, 32, 2, 40, 15, 0.0),
                                                                (1024, 0, 2, 700, 300, 560.0)])
def test_ja98_generic_shapes(gpu, nchan, freq_res, D, nframes, chunk, dm):
    """decode_mode = JA98 on the generic channeliser (the CLI default --nchan 512 among it): levels per window of 512 samples of
    the de-framed stream, windows counted through the samples carried from push to push; also under the generic dedispersion."""
    bw, fc = 32.0, 1254.0
    L = freq_res or 2 * nchan
    v = synth.make_vdif(nframes, seed=4300 + nchan, bw_mhz=bw, tone_frac=0.31, invalid_frac=0.02, fill_frac=0.02)
    cfg = PlanConfig(nchan=nchan, bw_mhz=[-bw], freq_mhz=[fc], freq_res=freq_res, tscrunch=D, out_nbit=-32, keep_bandpass=True,
                     decode_mode=_lib.DECODE_JA98, chunk_units=chunk, dm=dm, coherent=dm > 0)
    out = []
    with Plan(cfg) as pl:
        nf = (int(pl.geometry.nfilt_pos), int(pl.geometry.nfilt_neg))
        cf, fb = int(pl.chunk_frames), cfg.frame_bytes
        assert cf == chunk and int(pl.geometry.freq_res) == L
        for f0 in range(0, nframes, cf):
            n = min(cf, nframes - f0)
            pl.push([v[f0 * fb:(f0 + n) * fb]])
            out.append(pl.pull().copy())
        pl.flush()
        out.append(pl.pull().copy())
        rows = pl.view_rows(np.concatenate(out))
    kw = dict(freq_mhz=fc, bw_mhz=-bw, nchan=nchan, freq_res=L, tscrunch_factor=D, out_nbit=-32, keep_bandpass=True, dm=dm, coherent=dm > 0, nfilt=nf)
    ref = o.digifil(v, decode_mode="ja98", **kw)["data"].astype(np.float64)
    assert rows.shape[0] == ref.shape[0] > 0
    assert_rel(rows.reshape(ref.shape), ref, REL_TOL, f"JA98 generic nchan {nchan} L {L}")
    assert np.abs(o.digifil(v, **kw)["data"] / ref - 1).max() > 0.5      # not the static levels


@pytest.mark.parametrize("nchan,D,mode,name,dm,force_legacy", [(128, 16, _lib.POL_P0, "P0", 0.0, False), (64, 8, _lib.POL_I2, "I2", 0.0, False),
                                                               (128, 16, _lib.POL_I, "I", 0.0, True), (128, 16, _lib.POL_COHERENCE, "coherence", 560.0, False)])
def test_ja98_round1_kernels(gpu, monkeypatch, nchan, D, mode, name, dm, force_legacy):
    """decode_mode = JA98 where the round-2 column kernel does not apply (the other detection products, dedispersion,
    B2F_PATH=legacy): the round-1 column kernel reads the levels per window of 512 samples of the de-framed stream."""
    if force_legacy:
        monkeypatch.setenv("B2F_PATH", "legacy")
    bw, fc = 32.0, 1254.0
    cfg = PlanConfig(nchan=nchan, bw_mhz=[-bw], freq_mhz=[fc], tscrunch=D, pol_mode=mode, out_nbit=-32, keep_bandpass=True,
                     decode_mode=_lib.DECODE_JA98, dm=dm, coherent=dm > 0, chunk_units=8 if dm > 0 else 0)
    out = []
    with Plan(cfg) as pl:
        assert pl.path == 0
        nf = (int(pl.geometry.nfilt_pos), int(pl.geometry.nfilt_neg))
        cf, fb = int(pl.chunk_frames), cfg.frame_bytes
        nframes = cf + (cf // 2 if dm > 0 else 0)
        v = synth.make_vdif(nframes, seed=4400 + nchan, bw_mhz=bw, tone_frac=0.31, rho=0.3, invalid_frac=0.01, fill_frac=0.01)
        for f0 in range(0, nframes, cf):
            n = min(cf, nframes - f0)
            pl.push([v[f0 * fb:(f0 + n) * fb]])
            out.append(pl.pull().copy())
        pl.flush()
        out.append(pl.pull().copy())
        rows = pl.view_rows(np.concatenate(out))
    ref = o.digifil(v, freq_mhz=fc, bw_mhz=-bw, nchan=nchan, tscrunch_factor=D, pol_mode=name, out_nbit=-32, keep_bandpass=True,
                    decode_mode="ja98", dm=dm, coherent=dm > 0, nfilt=nf)["data"].astype(np.float64)
    assert rows.shape[0] == ref.shape[0] > 0
    power = ref[:, 0] + ref[:, 1] if name == "coherence" else None
    assert_rel(rows.reshape(ref.shape), ref, REL_TOL, f"JA98 round-1 kernels {name}", power=power)


def test_ja98_is_a_2bit_unpacker(gpu):
    with pytest.raises(_lib.B2FError) as e:
        Plan(PlanConfig(nchan=128, bw_mhz=[-32.0], in_nbit=8, decode_mode=_lib.DECODE_JA98))
    assert e.value.code == _lib.EINVAL


def test_8bit_offset_128(gpu):
    nchan, bw, D = 128, 32.0, 16
    with Plan(PlanConfig(nchan=nchan, bw_mhz=[-bw], in_nbit=8)) as pl:
        nfr = int(pl.chunk_frames)
    v = synth.make_vdif(nfr, seed=72, bw_mhz=bw, nbit=8)
    cfg = PlanConfig(nchan=nchan, bw_mhz=[-bw], tscrunch=D, in_nbit=8, out_nbit=-32, keep_bandpass=True, in8_offset_mode=1)
    rows, _ = _run(cfg, v)
    ref = o.digifil(v, freq_mhz=1400.0, bw_mhz=-bw, nchan=nchan, tscrunch_factor=D, in_nbit=8, out_nbit=-32, keep_bandpass=True,
                    offset8=128.0)["data"].astype(np.float64)
    assert_rel(rows.reshape(ref.shape), ref, REL_TOL, "8-bit, code - 128")


def test_digi_sigma_and_normalised_transforms(gpu):
    nchan, bw, D = 128, 32.0, 16
    v = synth.make_vdif(1024, seed=73, bw_mhz=bw, tone_frac=0.2)
    rows, _ = _run(PlanConfig(nchan=nchan, bw_mhz=[-bw], tscrunch=D, rescale_interval_s=0.2, digi_sigma=3.0), v)
    ref = o.digifil(v, freq_mhz=1400.0, bw_mhz=-bw, nchan=nchan, tscrunch_factor=D, rescale_interval_s=0.2, digi_sigma=3.0)["data"]
    d = np.abs(rows.astype(int) - ref.reshape(rows.shape).astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3
    assert rows.std() > 1.7 * 21.25                            # sigma now maps to 42.5 counts
    rows, _ = _run(PlanConfig(nchan=nchan, bw_mhz=[-bw], tscrunch=D, out_nbit=-32, keep_bandpass=True, fft_normalised=True), v)
    ref = o.digifil(v, freq_mhz=1400.0, bw_mhz=-bw, nchan=nchan, tscrunch_factor=D, out_nbit=-32, keep_bandpass=True,
                    fft_normalised=True)["data"].astype(np.float64)
    assert_rel(rows.reshape(ref.shape), ref, REL_TOL, "normalised transforms")
    assert 1.0 < ref.mean() / D < 20.0                         # about the input power per sample (sigma^2 of 2 pols), not 1e8


def test_running_rescale(gpu):
    """rescale_mode = RUNNING (digifil without -c): every 0.2 s interval is scaled with its own mean / sigma; pushes of
    0.256 s so that intervals straddle pushes and held rows have to be moved to the front of the row buffer."""
    nchan, bw, D = 128, 32.0, 16
    v = synth.make_vdif(5 * 1024 + 300, seed=74, bw_mhz=bw, tone_frac=0.4)
    cfg = PlanConfig(nchan=nchan, bw_mhz=[-bw], tscrunch=D, rescale_interval_s=0.2, rescale_mode=_lib.RESCALE_RUNNING, chunk_units=1)
    rows, _ = _run(cfg, v, pulls_until_empty=True)
    ref = o.digifil(v, freq_mhz=1400.0, bw_mhz=-bw, nchan=nchan, tscrunch_factor=D, rescale_interval_s=0.2,
                    rescale_mode="running")["data"]
    assert rows.shape[0] == ref.shape[0]
    d = np.abs(rows.astype(int) - ref.reshape(rows.shape).astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3
    const = o.digifil(v, freq_mhz=1400.0, bw_mhz=-bw, nchan=nchan, tscrunch_factor=D, rescale_interval_s=0.2)["data"]
    assert (const.reshape(rows.shape) != ref.reshape(rows.shape)).mean() > 0.01
