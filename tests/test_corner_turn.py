"""Corner turn: raw multi-BBC VDIF -> per-IF samples on the GPU,
replacing jive5ab's spif2file split (/root/reference/spif2file.sh)."""
import numpy as np
import pytest

from oracle import digifil_oracle as o
from frb_baseband_b200 import spif, synth
from frb_baseband_b200.plan import Plan, PlanConfig
from helpers import REL_TOL, assert_rel


def test_recipes_follow_spif2file():
    W, g = spif.recipe_for_mode("VDIF_8000-2048-16-2", 8)                     # spif2file.sh:40-42
    assert W == 32 and g[0] == [16, 17, 24, 25] and g[1] == [0, 1, 8, 9] and g[7] == [6, 7, 14, 15]
    W, g = spif.recipe_for_mode("VDIF_8000-4096-32-2", 16)                    # :32-34
    assert W == 64 and g[0] == [16, 17, 48, 49] and g[15] == [14, 15, 46, 47]
    W, g = spif.recipe_for_mode("VDIF_8000-1024-16-2", 8)                     # :44-47 (mirrored pairing)
    assert W == 32 and g[0] == [24, 25, 16, 17]
    W, g = spif.recipe_for_mode("VDIF_8000-1024-8-2", 4)                      # :54-56
    assert W == 16 and g == [[8, 9, 12, 13], [0, 1, 4, 5], [10, 11, 14, 15], [2, 3, 6, 7]]
    _, f = spif.recipe_for_mode("VDIF_8000-2048-16-2", 8, flip_if=True)       # :117-131: neighbours swap
    assert f[0] == [0, 1, 8, 9] and f[1] == [16, 17, 24, 25] and f[6] == [6, 7, 14, 15]
    assert spif.mode_string(8000, 2048, 16, 2) == "VDIF_8000-2048-16-2"       # base2fil.sh:308-318
    with pytest.raises(ValueError):
        spif.recipe_for_mode("VDIF_8000-77-3-2", 1)
    # every recipe uses each input bit exactly once
    for mode, nif in (("VDIF_8000-2048-16-2", 8), ("VDIF_8000-4096-32-2", 16), ("VDIF_8000-1024-8-2", 4)):
        W, g = spif.recipe_for_mode(mode, nif)
        assert sorted(b for grp in g for b in grp) == list(range(W))


@pytest.mark.parametrize("mode,nif", [("VDIF_8000-2048-16-2", 8), ("VDIF_8000-4096-32-2", 16), ("VDIF_8000-1024-8-2", 4)])
def test_oracle_corner_turn_inverts_synthetic_raw(mode, nif):
    W, bits = spif.recipe_for_mode(mode, nif)
    rng = np.random.default_rng(5)
    spf = 8000 * 8 // W
    codes = rng.integers(0, 4, size=(nif, 2, 3 * spf), dtype=np.uint8)
    raw = synth.make_raw_vdif(codes, W, bits, bw_mhz=32.0)
    x = o.corner_turn(raw, W, bits)
    assert x.shape == (nif, 2, 3 * spf)
    assert np.array_equal(x, o.LEVELS_2BIT[codes])
    raw2 = raw.copy().reshape(3, 8032)
    raw2[1, 3] |= 0x80                                                       # invalid bit of frame 1
    x2 = o.corner_turn(raw2.reshape(-1), W, bits).reshape(nif, 2, 3, spf)
    assert np.all(x2[:, :, 1] == 0) and np.array_equal(x2[:, :, 0], x.reshape(nif, 2, 3, spf)[:, :, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("mode,nif,bw", [("VDIF_8000-2048-16-2", 8, 32.0), ("VDIF_8000-4096-32-2", 16, 32.0),
                                         ("VDIF_8000-1024-8-2", 4, 32.0)])
def test_gpu_corner_turn_equals_split_path(gpu, mode, nif, bw):
    """One raw stream in, spliced filterbank out == oracle(corner turn -> digifil per IF -> splice)."""
    nchan, D = 32, 32
    W, bits = spif.recipe_for_mode(mode, nif)
    bws = [bw if i % 2 == 0 else -bw for i in range(1, nif + 1)]
    freqs = [1300.0 + (i - 1) * bw for i in range(1, nif + 1)]
    cfg = PlanConfig(nchan=nchan, bw_mhz=bws, freq_mhz=freqs, tscrunch=D, out_nbit=-32, keep_bandpass=True,
                     raw_word_bits=W, raw_bits=bits, frame_bytes=8032)
    with Plan(cfg) as pl:
        nfr = int(pl.chunk_frames)
        spf = 8000 * 8 // W
        rng = np.random.default_rng(17)
        codes = synth.quantise_2bit(rng.standard_normal((nif, 2, nfr * spf)) + 0.4 * np.cos(0.9 * np.arange(nfr * spf)))
        raw = synth.make_raw_vdif(codes, W, bits, bw_mhz=bw).reshape(nfr, 8032)
        raw[5, 3] |= 0x80                                                    # an invalid frame
        raw[9, 32 + 400:32 + 440].view("<u4")[:] = 0x11223344                # a run of fill words
        pl.push([raw.reshape(-1)])
        rows = pl.view_rows(pl.pull())
        c = pl.counters()
    assert c["frames_invalid"] == 1 and c["frames_with_fill"] == 1 and c["frames_ok"] == nfr - 2
    x = o.corner_turn(raw.reshape(-1), W, bits)
    parts = []
    for i in sorted(range(nif), key=lambda k: -freqs[k]):                     # splice: highest frequency first
        parts.append(o.digifil(raw.reshape(-1)[:8032], freq_mhz=freqs[i], bw_mhz=bws[i], nchan=nchan, tscrunch_factor=D,
                               out_nbit=-32, keep_bandpass=True, x=x[i]))
    ref = o.splice(parts)["data"]
    assert rows.shape == ref.shape
    assert_rel(rows.reshape(rows.shape[0], 1, -1), ref.reshape(ref.shape[0], 1, -1).astype(np.float64), REL_TOL, "corner turn " + mode)
