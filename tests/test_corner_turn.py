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


def test_mark5b_recipes_fold_swap_sign_mag():
    """spif2file.sh:79-93: `swap_sign_mag+<recipe>`.  The host folds the swap into the bit positions (b -> b ^ 1); the
    oracle does the swap literally on the words and then applies the recipe as written."""
    W, g = spif.recipe_for_mode("MARK5B-1024-16-2", 8)                        # :79-81
    Wv, gv = spif.recipe_for_mode("VDIF_8000-2048-16-2", 8)
    assert W == Wv == 32 and g == [[b ^ 1 for b in grp] for grp in gv] and g[0] == [17, 16, 25, 24]
    W16, g16 = spif.recipe_for_mode("MARK5B-1024-8-2", 4)                     # :83-85
    assert W16 == 16 and g16[0] == [9, 8, 13, 12]
    assert spif.parse_recipe("swap_sign_mag+16>[8,9,12,13][0,1,4,5][10,11,14,15][2,3,6,7]:0-3") == (16, g16)
    assert spif.frame_geometry("MARK5B-2048-16-2") == (10016, 16, 1)          # :105-108
    assert spif.frame_geometry("VDIF_8000-2048-16-2") == (8032, 32, 0) and spif.frame_geometry("VDIF_1000-1024-16-2")[0] == 1032
    _, f = spif.recipe_for_mode("MARK5B-2048-16-2", 8, flip_if=True)
    assert f[0] == g[1] and f[1] == g[0]
    W64, g64 = spif.recipe_for_mode("MARK5B-2048-32-2", 16)                   # :91-93
    assert W64 == 64 and g64[0] == [17, 16, 49, 48] and sorted(b for grp in g64 for b in grp) == list(range(64))
    with pytest.raises(ValueError):
        spif.recipe_for_mode("MARK5B-512-16-2", 8)
    rng = np.random.default_rng(6)
    spf = 10000 * 8 // W
    codes = rng.integers(0, 4, size=(8, 2, 3 * spf), dtype=np.uint8)
    raw = synth.make_raw_mark5b(codes, W, g, bw_mhz=16.0, sec0=86399)
    assert raw.size == 3 * 10016
    hw = raw.reshape(3, 10016)[:, :16].copy().view("<u4")
    assert np.all(hw[:, 0] == 0xABADDEED) and list(hw[:, 1]) == [0, 1, 2] and hw[0, 2] == 0x00086399   # BCD JJJSSSSS
    x = o.corner_turn(raw, W, g, frame_bytes=10016, header_bytes=16, mark5b=True)
    assert np.array_equal(x, o.LEVELS_2BIT[codes])
    x_lit = o.corner_turn(raw, W, gv, frame_bytes=10016, header_bytes=16, mark5b=True, swap_sign_mag=True)
    assert np.array_equal(x_lit, x)
    raw2 = raw.copy().reshape(3, 10016)
    raw2[1, 0] ^= 1                                                          # broken sync word
    raw2[2, 5] |= 0x80                                                       # test-vector bit (bit 15 of word 1)
    x2 = o.corner_turn(raw2.reshape(-1), W, g, frame_bytes=10016, header_bytes=16, mark5b=True).reshape(8, 2, 3, spf)
    assert np.all(x2[:, :, 1:] == 0) and np.array_equal(x2[:, :, 0], x.reshape(8, 2, 3, spf)[:, :, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("mode,nif,bw,by_header", [("MARK5B-1024-16-2", 8, 16.0, False), ("MARK5B-1024-8-2", 4, 32.0, False),
                                                   ("MARK5B-2048-16-2", 8, 32.0, True), ("MARK5B-2048-32-2", 16, 16.0, False),
                                                   ("MARK5B-2048-32-2", 16, 16.0, True)])
def test_gpu_mark5b_corner_turn(gpu, mode, nif, bw, by_header):
    """Mark5B disk frames in, spliced filterbank out == oracle(swap_sign_mag -> recipe as written -> digifil per IF -> splice).
    by_header: 17 frames are missing from the recording and the push crosses a second boundary; frames are placed by the BCD
    time code and the frame number within the second, the gap is zero-filled."""
    nchan, D, nfr = 32, 32, 1024
    W, bits = spif.recipe_for_mode(mode, nif)
    _, written = spif.recipe_for_mode({32: "VDIF_8000-4096-32-2", 16: "VDIF_8000-2048-16-2", 8: "VDIF_8000-1024-8-2"}[2 * nif], nif)
    fb, hb, fmt = spif.frame_geometry(mode)
    bws = [bw if i % 2 == 0 else -bw for i in range(1, nif + 1)]
    freqs = [1300.0 + (i - 1) * bw for i in range(1, nif + 1)]
    cfg = PlanConfig(nchan=nchan, bw_mhz=bws, freq_mhz=freqs, tscrunch=D, out_nbit=-32, keep_bandpass=True,
                     raw_word_bits=W, raw_bits=bits, raw_format=fmt, frame_bytes=fb, header_bytes=hb, frame_time_mode=int(by_header))
    spf = 10000 * 8 // W
    fps = int(round(2 * bw * 1e6 / spf))
    rng = np.random.default_rng(19)
    codes = synth.quantise_2bit(rng.standard_normal((nif, 2, nfr * spf)) + 0.4 * np.cos(0.9 * np.arange(nfr * spf)))
    raw = synth.make_raw_mark5b(codes, W, bits, bw_mhz=bw, sec0=86399, frame0=fps - 500).reshape(nfr, fb)   # crosses midnight too
    raw[5, 0] ^= 1                                                           # a frame without the sync word
    raw[7, 5] |= 0x80                                                        # a test-vector frame
    raw[9, hb + 400:hb + 440].view("<u4")[:] = 0x11223344                    # a run of fill words
    sent = raw
    if by_header:
        lost = np.arange(600, 617)
        filler = np.repeat(raw[:1], lost.size, axis=0)
        filler[:, 8:12] = np.frombuffer(np.uint32(0x50000000).tobytes(), np.uint8)       # day 500: far outside the push -> dropped
        sent = np.concatenate([np.delete(raw, lost, axis=0), filler])
        raw[lost, 5] |= 0x80                                                 # oracle: the missing frames carry no data
    with Plan(cfg) as pl:
        assert int(pl.chunk_frames) >= nfr and int(pl.geometry.samples_per_frame) == spf
        pl.push([sent.reshape(-1)])
        rows = pl.view_rows(pl.pull())
        c = pl.counters()
    assert c["frames_badhdr"] == 1 and c["frames_invalid"] == 1 and c["frames_with_fill"] == 1
    assert c["frames_ok"] == nfr - 3 - (17 if by_header else 0) and c["slots_missing"] == (17 * nif if by_header else 0)      # counted per IF stream
    assert c["frames_dropped"] == (17 if by_header else 0)
    x = o.corner_turn(raw.reshape(-1), W, written, frame_bytes=fb, header_bytes=hb, mark5b=True, swap_sign_mag=True)
    parts = []
    for i in sorted(range(nif), key=lambda k: -freqs[k]):
        parts.append(o.digifil(np.zeros(8032, np.uint8), freq_mhz=freqs[i], bw_mhz=bws[i], nchan=nchan, tscrunch_factor=D,
                               out_nbit=-32, keep_bandpass=True, x=x[i], frame_bytes=8032))   # x given: the bytes are not decoded
    ref = o.splice(parts)["data"]
    assert rows.shape == ref.shape and rows.shape[0] > 0
    assert_rel(rows.reshape(rows.shape[0], 1, -1), ref.reshape(ref.shape[0], 1, -1).astype(np.float64), REL_TOL, "corner turn " + mode)


@pytest.mark.gpu
def test_gpu_mark5b_file_to_filterbank(gpu, tmp_path):
    """base2fil.run_scan_raw on a Mark5B recording: same rows as the streaming calls, start time from the BCD time code."""
    from frb_baseband_b200 import base2fil, sigproc
    mode, nif, bw, nchan, D = "MARK5B-1024-16-2", 8, 16.0, 8, 32
    W, bits = spif.recipe_for_mode(mode, nif)
    fb, hb, fmt = spif.frame_geometry(mode)
    bws = [bw if i % 2 == 0 else -bw for i in range(1, nif + 1)]
    freqs = [1300.0 + (i - 1) * bw for i in range(1, nif + 1)]
    cfg = PlanConfig(nchan=nchan, bw_mhz=bws, freq_mhz=freqs, tscrunch=D, out_nbit=-32, keep_bandpass=True,
                     raw_word_bits=W, raw_bits=bits, raw_format=fmt, frame_bytes=fb, header_bytes=hb)
    with Plan(cfg) as pl:
        nfr = int(pl.geometry.unit_frames)
        assert nfr * fb < 64e6
        rng = np.random.default_rng(23)
        codes = rng.integers(0, 4, size=(nif, 2, nfr * 2500), dtype=np.uint8)
        raw = synth.make_raw_mark5b(codes, W, bits, bw_mhz=bw, mjd=60123, sec0=3600)
        rows = []
        cf = int(pl.chunk_frames)
        for f0 in range(0, nfr, cf):
            pl.push([raw[f0 * fb:(f0 + min(cf, nfr - f0)) * fb]])
            rows.append(pl.pull().copy())
        pl.flush()
        rows.append(pl.pull().copy())
        rows = np.concatenate(rows)
    path, out = tmp_path / "scan.m5b", tmp_path / "scan.fil"
    raw.tofile(path)
    base2fil.run_scan_raw(str(path), str(out), mode=mode, nif=nif, bw=bw, freq_lsb0=1300.0, nchan=nchan, tscrunch=D,
                          nbit=-32, keep_bandpass=True, verbose=False)
    hdr, data = sigproc.read_fil(str(out))
    assert hdr.nchans == nif * nchan and hdr.nbits == 32
    # JJJ = 123: the thousands come from the clock (latest such day that is not in the future)
    assert abs(hdr.tstart % 1000 - (123 + 3600 / 86400.0)) < 1e-9 and hdr.tstart > 1000
    assert np.array_equal(np.asarray(data).reshape(-1).view(np.uint8), rows.reshape(-1).view(np.uint8))


def test_one_bit_recipe_and_oracle():
    """VDIF_8000-1024-16-1 (spif2file.sh:58-61): 16 BBC channels of 1-bit samples, two recipe bits per IF."""
    W, g = spif.recipe_for_mode("VDIF_8000-1024-16-1", 8)
    assert W == 16 and g[0] == [8, 12] and g[7] == [3, 7] and sorted(b for grp in g for b in grp) == list(range(16))
    _, f = spif.recipe_for_mode("VDIF_8000-1024-16-1", 8, flip_if=True)
    assert f[0] == [0, 4] and f[1] == [8, 12]
    rng = np.random.default_rng(8)
    codes = rng.integers(0, 2, size=(8, 2, 2 * 4000), dtype=np.uint8)
    raw = synth.make_raw_vdif(codes, W, g, bw_mhz=32.0)
    assert raw.size == 2 * 8032 and (int(raw[15]) >> 2) & 31 == 0             # header: bits per sample - 1 = 0
    x = o.corner_turn(raw, W, g)
    assert np.array_equal(x, 2.0 * codes - 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,nif,bw", [("VDIF_8000-2048-16-2", 8, 32.0), ("VDIF_8000-4096-32-2", 16, 32.0),
                                         ("VDIF_8000-1024-8-2", 4, 32.0), ("VDIF_8000-1024-16-1", 8, 32.0)])
def test_gpu_corner_turn_equals_split_path(gpu, mode, nif, bw):
    """One raw stream in, spliced filterbank out == oracle(corner turn -> digifil per IF -> splice)."""
    nchan, D = 32, 32
    W, bits = spif.recipe_for_mode(mode, nif)
    bws = [bw if i % 2 == 0 else -bw for i in range(1, nif + 1)]
    freqs = [1300.0 + (i - 1) * bw for i in range(1, nif + 1)]
    cfg = PlanConfig(nchan=nchan, bw_mhz=bws, freq_mhz=freqs, tscrunch=D, out_nbit=-32, keep_bandpass=True,
                     raw_word_bits=W, raw_bits=bits, frame_bytes=8032, in_nbit=len(bits[0]) // 2)
    with Plan(cfg) as pl:
        nfr = int(pl.chunk_frames)
        spf = 8000 * 8 // W
        rng = np.random.default_rng(17)
        sig = rng.standard_normal((nif, 2, nfr * spf)) + 0.4 * np.cos(0.9 * np.arange(nfr * spf))
        codes = synth.quantise_2bit(sig) if len(bits[0]) == 4 else (sig > 0).astype(np.uint8)
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
