"""The reference-facing entry points on a GPU: process_vdif CLI (regular file and FIFO) and the
scan driver, against the oracle."""
import os
import threading

import numpy as np
import pytest

from oracle import digifil_oracle as o
from frb_baseband_b200 import base2fil, process_vdif, sigproc, synth

pytestmark = pytest.mark.gpu


def _write_vdif(path, nframes, seed, bw, **kw):
    v = synth.make_vdif(nframes, seed=seed, bw_mhz=bw, **kw)
    v.tofile(path)
    return v


def test_process_vdif_cli_regular_file(gpu, tmp_path):
    vd = tmp_path / "ek_ef_no0001_IF3.vdif"
    v = _write_vdif(str(vd), 4000 + 1024, 5, 32.0, tone_frac=0.4)
    out = tmp_path / "fil"
    out.mkdir()
    rc = process_vdif.main(["R3", "--ra", "01:58:00.75", "--dec=65:43:00.31", str(vd), "-f", "1318.0", "-b", "32", "-l",
                            "--nchan", "128", "--nsec", "1", "--start", "0.256", "--force", "-t", "effelsberg", "--pol", "2",
                            "--nthreads", "1", "--tscrunch", "16", "--fil_out_dir", str(out), "--nbit=8"])
    assert rc == 0
    fil = out / "ek_ef_no0001_IF3.vdif_pol2.fil"
    assert fil.exists() and (tmp_path / "ek_ef_no0001_IF3.vdif_pol2.hdr").exists()
    h, d = sigproc.read_fil(str(fil))
    ref = o.digifil(v, freq_mhz=1318.0, bw_mhz=-32.0, nchan=128, tscrunch_factor=16, start_s=0.256, nsec=1.0)
    assert (h.nchans, h.nifs, h.nbits, h.source_name, h.telescope_id) == (128, 1, 8, "R3", 8)
    assert h.tsamp == pytest.approx(64e-6) and h.foff == pytest.approx(-0.25) and h.fch1 == pytest.approx(ref["fch1"])
    assert h.tstart == pytest.approx(ref["tstart"], abs=1e-9)
    assert d.shape == ref["data"].shape
    assert np.abs(d.astype(int) - ref["data"].astype(int)).max() <= 1


def test_process_vdif_writes_into_existing_fifo(gpu, tmp_path):
    """base2fil.sh:348-349 makes the FIFO first; splice reads it (here: a reader thread)."""
    vd = tmp_path / "ek_ef_no0001_IF2.vdif"
    v = _write_vdif(str(vd), 1024, 6, 16.0)
    fifo = tmp_path / "ek_ef_no0001_IF2.vdif_pol2.fil"
    os.mkfifo(fifo)
    got = {}

    def reader():
        import fcntl
        with open(fifo, "rb") as f:
            first = f.read(1)                       # the writer has opened (and resized) the pipe by now
            got["pipe_size"] = fcntl.fcntl(f.fileno(), getattr(fcntl, "F_GETPIPE_SZ", 1032))
            got["raw"] = first + f.read()

    t = threading.Thread(target=reader)
    t.start()
    process_vdif.main(["SRC", "--ra", "00:00:00", "--dec", "00:00:00", str(vd), "-f", "1400", "-b", "16", "-u", "--nchan", "32",
                       "--nsec", "10", "--start", "0", "--force", "--tscrunch", "32", "--fil_out_dir", str(tmp_path)])
    t.join(60)
    assert not t.is_alive() and os.path.exists(fifo)
    assert got["pipe_size"] == 1048576              # setfifo.perl:10 (F_SETPIPE_SZ, 1 MiB) for every FIFO of the splice list
    h, off = sigproc.read_header(got["raw"])
    d = np.frombuffer(got["raw"], np.uint8, offset=off).reshape(-1, 32)
    ref = o.digifil(v, freq_mhz=1400.0, bw_mhz=16.0, nchan=32, tscrunch_factor=32)
    assert np.abs(d.astype(int) - ref["data"].reshape(d.shape).astype(int)).max() <= 1


def test_base2fil_conf_to_spliced_file(gpu, tmp_path):
    """frb.conf -> final IFall file, 4 IFs x 16 MHz (BASELINE config 1, 512 Mbps reading)."""
    odd, even, outd = tmp_path / "s0" / "c1test", tmp_path / "s1" / "c1test", tmp_path / "out"
    for d in (odd, even):
        d.mkdir(parents=True)
    nif, bw = 4, 16.0
    vd = {}
    for i in range(1, nif + 1):
        p = (odd if i % 2 else even) / f"c1test_o8_no0007_IF{i}.vdif"
        vd[i] = _write_vdif(str(p), 2000, synth.config_seed(1, i), bw, tone_frac=0.15 * i)
    c = tmp_path / "frb.conf"
    c.write_text(f"experiment=c1test\ntarget=\"B0329+54 --ra 03:32:59.4 --dec +54:34:43.3\"\nscans=( 7 )\nskips=( 0 )\n"
                 f"lengths=( 1 )\nscannames=( 007 )\nbw=16\nnif={nif}\nfreqLSB_0=1300.0\nstation=onsala85\nnchan=32\n"
                 f"tscrunch=32\nworkdir_odd_base={tmp_path}/s0\nworkdir_even_base={tmp_path}/s1\noutdir_base={tmp_path}/out\n"
                 f"submit2fetch=1\nflagFile=/flags/o8.flag_1284-1348MHz_128chan\n")
    sent = []
    done = base2fil.base2fil(str(c), send=lambda argv: sent.append(argv) or 0)
    # base2fil.sh:420-435: keepVDIF=0 (default) removes the split files, then "<fil> <flag>" goes to stage01_queue
    assert not any(os.path.exists(str((odd if i % 2 else even) / f"c1test_o8_no0007_IF{i}.vdif")) for i in range(1, nif + 1))
    assert len(sent) == 1 and sent[0][-4:] == ["-q", "stage01_queue", "-m", done[0] + " /flags/o8.flag_1284-1348MHz_128chan"]
    assert done == [str(tmp_path / "out" / "c1test" / "c1test_o8_no0007_IFall_vdif_pol2.fil")]
    h, d = sigproc.read_fil(done[0])
    ref = o.base2fil(vd, nif=nif, freq_lsb0=1300.0, bw=bw, nchan=32, tscrunch_factor=32, nsec=1.0)
    assert h.nchans == 128 and h.fch1 == pytest.approx(ref["fch1"]) and h.source_name == "B0329+54"
    assert d.shape[0] == ref["data"].shape[0] and d.shape[2] == 128
    assert np.abs(d[:, 0, :].astype(int) - ref["data"].astype(int)).max() <= 1


def test_base2fil_raw_conf_to_spliced_file(gpu, tmp_path):
    """Mode C from a frb.conf: the raw multi-BBC recording is the only input; window, recipe and frequency plan as
    base2fil.sh / spif2file.sh derive them.  Checked against oracle(corner turn -> digifil per IF -> splice)."""
    from frb_baseband_b200 import spif
    nif, bw, nchan, D = 4, 2.0, 8, 4
    vbs = tmp_path / "vbs" / "rawtest"
    vbs.mkdir(parents=True)
    W, bits = spif.recipe_for_mode("VDIF_8000-64-8-2", nif, flip_if=True)
    spf, fps, nfr = 8000 * 8 // W, 1000, 3000
    rng = np.random.default_rng(29)
    codes = synth.quantise_2bit(rng.standard_normal((nif, 2, nfr * spf)) + 0.5 * np.cos(0.7 * np.arange(nfr * spf)))
    raw = synth.make_raw_vdif(codes, W, bits, bw_mhz=bw, sec0=5, frame0=37)
    raw.tofile(vbs / "rawtest_ef_no0012")
    c = tmp_path / "frb.conf"
    c.write_text(f"experiment=rawtest\ntarget=\"R3 --ra 01:58:00.75 --dec +65:43:00.3\"\nscans=( 012 )\nskips=( 1 )\nlengths=( 1 )\n"
                 f"scannames=( 012 )\nbw={bw}\nnif={nif}\nfreqLSB_0=1300.0\nstation=effelsberg\nnchan={nchan}\ntscrunch={D}\nflipIF=1\n"
                 f"nbit=-32\nkeepBP=1\nvbsdir_base={tmp_path}/vbs\noutdir_base={tmp_path}/out\n")
    done = base2fil.base2fil_raw(str(c), verbose=False)
    assert done == [str(tmp_path / "out" / "rawtest" / "rawtest_ef_no0012_IFall_vdif_pol2.fil")]
    h, d = sigproc.read_fil(done[0])
    f0 = (fps - 37 - 1) + fps                                             # spif2file.sh:144-149
    x = o.corner_turn(raw, W, bits)[:, :, f0 * spf:(f0 + fps) * spf]
    freqs = [1300.0 + (i - 1) * bw for i in range(1, nif + 1)]
    bws = [bw if i % 2 == 0 else -bw for i in range(1, nif + 1)]
    parts = [o.digifil(np.zeros(8032, np.uint8), freq_mhz=freqs[i], bw_mhz=bws[i], nchan=nchan, tscrunch_factor=D, out_nbit=-32,
                       keep_bandpass=True, x=x[i], frame_bytes=8032) for i in sorted(range(nif), key=lambda k: -freqs[k])]
    ref = o.splice(parts)["data"]
    assert h.nchans == nif * nchan and h.nbits == 32 and h.source_name == "R3" and d.shape[0] == ref.shape[0] > 0
    assert h.tstart == pytest.approx(vdif_mjd(raw[f0 * 8032:f0 * 8032 + 32], fps), abs=1e-12)
    from helpers import REL_TOL, assert_rel
    assert_rel(d.reshape(d.shape[0], 1, -1), ref.reshape(ref.shape[0], 1, -1).astype(np.float64), REL_TOL, "mode C from a conf")


def test_native_runner_equals_streaming_calls(gpu, tmp_path):
    """b2f_run_scan (threaded readers, pinned ring, pipelined push/pull) writes exactly the rows the caller gets
    from b2f_push/b2f_pull chunk by chunk, for a -S/-T window that ends inside a chunk, files of unequal length,
    and every ring depth."""
    from frb_baseband_b200.plan import Plan, PlanConfig
    from helpers import run_plan
    nif, bw, fb = 2, 16.0, 8032
    nfr = [4 * 1000 + 650, 4 * 1000 + 900]
    paths, vd = [], []
    for i in range(nif):
        p = tmp_path / f"x_IF{i + 1}.vdif"
        vd.append(_write_vdif(str(p), nfr[i], 40 + i, bw, tone_frac=0.2 + 0.3 * i, invalid_frac=0.002))
        paths.append(str(p))
    bws, freqs = [-bw, bw], [1300.0, 1316.0]
    start, nsec = 0.25, 1.7                                   # 2000 frames/s at 16 MHz: frames 500 .. 3900
    f0, n = 500, 3400
    ref, info = run_plan([v[f0 * fb:(f0 + n) * fb] for v in vd], nchan=32, bw=bws, freq=freqs, tscrunch=32, interval=0.5)
    for ring in (2, 5, 0):
        out = tmp_path / f"o{ring}.fil"
        with Plan(PlanConfig(nchan=32, bw_mhz=bws, freq_mhz=freqs, tscrunch=32, rescale_interval_s=0.5)) as pl:
            r = pl.run_scan(paths, str(out), start_s=start, nsec=nsec, source_name="SRC", ring=ring)
            r2 = pl.run_scan(paths, str(tmp_path / "again.fil"), start_s=start, nsec=nsec, source_name="SRC", ring=ring)
        h, d = sigproc.read_fil(str(out))
        assert r["frames_per_if"] == n and r["rows"] == len(ref) == d.shape[0]
        assert np.array_equal(d.reshape(len(ref), -1), ref)
        assert h.nchans == 64 and h.fch1 == pytest.approx(1316.0 + 8 - 0.25) and h.foff == pytest.approx(-0.5)
        assert h.tstart == pytest.approx(vdif_mjd(vd[0][f0 * fb:f0 * fb + 32], 2000))
        assert r["counters"]["frames_invalid"] == info["counters"]["frames_invalid"] > 0
        assert open(out, "rb").read() == open(tmp_path / "again.fil", "rb").read()      # a plan serves consecutive scans
    # to the end of the shortest file when -T is not given
    with Plan(PlanConfig(nchan=32, bw_mhz=bws, freq_mhz=freqs, tscrunch=32, rescale_interval_s=0.5)) as pl:
        r = pl.run_scan(paths, str(tmp_path / "all.fil"))
        assert r["frames_per_if"] == min(nfr)
        with pytest.raises(Exception) as e:
            pl.run_scan(paths[:1], str(tmp_path / "bad.fil"))
        assert "expects 2 input" in str(e.value)
        with pytest.raises(Exception) as e:
            pl.run_scan([paths[0], str(tmp_path / "missing.vdif")], str(tmp_path / "bad.fil"))
        assert "cannot open" in str(e.value)


def vdif_mjd(head, fps):
    from frb_baseband_b200 import vdif
    return vdif.frame_mjd(vdif.parse_header(bytes(head)), fps)


@pytest.mark.parametrize("kind", ["tuned", "dedisp", "generic", "float_bandpass"])
def test_scan_in_time_parts_is_bit_identical(gpu, tmp_path, kind):
    """SURVEY 8e time sharding: a scan processed as n consecutive parts -- each with the statistics of the scan's
    first interval preset and writing at its final file offset -- gives the file a single run gives, byte for byte
    (parts run one after the other here; on several GPUs they run concurrently, dist.run_scan_time_sharded)."""
    from frb_baseband_b200.plan import Plan, PlanConfig
    bw, nfr = 32.0, 3 * 1024 + 700
    kw = dict(nchan=128, bw_mhz=[-bw, bw], freq_mhz=[1300.0, 1332.0], tscrunch=16, rescale_interval_s=0.3)
    if kind == "dedisp":
        kw.update(dm=300.0, coherent=True)
    elif kind == "generic":
        kw.update(nchan=64, freq_res=256, tscrunch=8)
    elif kind == "float_bandpass":
        kw.update(out_nbit=-32, keep_bandpass=True, pol_mode=4)
    paths = []
    for i in range(2):
        p = tmp_path / f"t_IF{i + 1}.vdif"
        _write_vdif(str(p), nfr + 37 * i, 70 + i, bw, tone_frac=0.15 + 0.5 * i, rho=0.3)
        paths.append(str(p))
    whole = tmp_path / "whole.fil"
    with Plan(PlanConfig(**kw)) as pl:
        r = pl.run_scan(paths, str(whole), source_name="S")
        assert r["rows"] > 0
    ref = open(whole, "rb").read()
    for n in (2, 3):
        out = tmp_path / f"parts{n}.fil"
        with Plan(PlanConfig(**kw)) as pl:
            if kind != "float_bandpass":
                with pytest.raises(Exception) as e:
                    pl.run_scan(paths, str(out), source_name="S", part=(0, n))
                assert "b2f_set_rescale" in str(e.value)
                pl.run_scan(paths, None, stats_only=True)
                assert pl.counters()["rescale_frozen"] == 1
                pl.set_rescale(*pl.rescale())
            rows = 0
            for k in reversed(range(n)):                     # any order: every part knows its offset
                rows += pl.run_scan(paths, str(out), source_name="S", part=(k, n))["rows"]
        assert rows == r["rows"]
        assert open(out, "rb").read() == ref, f"{kind}: {n} parts differ from the single run"


def test_one_plan_two_scans_measures_each_scan(gpu, tmp_path):
    """A plan reused for a second, different scan through the time-sharded entry re-measures the rescale: a preset left
    from the first scan survives b2f_reset, and without clearing it the stats_only pass would stop after one chunk and
    hand out the first scan's mean/scale (dist.measure_first_interval is what run_scan_time_sharded calls on rank 0)."""
    from frb_baseband_b200 import dist
    from frb_baseband_b200.plan import Plan, PlanConfig
    bw, nfr = 32.0, 2 * 1024
    kw = dict(nchan=128, bw_mhz=[-bw], tscrunch=16, rescale_interval_s=0.2)
    pa, pb = tmp_path / "a.vdif", tmp_path / "b.vdif"
    _write_vdif(str(pa), nfr, 301, bw, tone_frac=0.2)
    _write_vdif(str(pb), nfr, 302, bw, tone_frac=0.7, tone_amp=1.5)
    with Plan(PlanConfig(**kw)) as fresh:
        want_b = dist.measure_first_interval(fresh, [str(pb)])
    with Plan(PlanConfig(**kw)) as pl:
        got_a = dist.measure_first_interval(pl, [str(pa)])
        pl.set_rescale(*got_a)                                   # what broadcast_rescale leaves behind on every rank
        pl.run_scan([str(pa)], str(tmp_path / "a.fil"), part=(0, 2))
        got_b = dist.measure_first_interval(pl, [str(pb)])       # second scan, same plan
        assert pl.counters()["rescale_preset"] == 0
    assert not np.allclose(got_a[0], got_b[0])
    assert np.array_equal(got_b[0], want_b[0]) and np.array_equal(got_b[1], want_b[1])


def test_edge_inputs_empty_short_and_widest(gpu, tmp_path):
    """Inputs at the edges of what base2fil can hand over: an empty split file (base2fil.sh:391-394 writes an
    empty .fil for it; the runner writes a header-only one), a file shorter than one FFT block (no samples), a
    window that starts beyond the end, and the widest plan (32 IFs, the B2F_MAX_IF of include/b2f.h)."""
    from frb_baseband_b200.plan import Plan, PlanConfig
    bw = 16.0
    empty, short = tmp_path / "e_IF1.vdif", tmp_path / "s_IF1.vdif"
    empty.write_bytes(b"")
    _write_vdif(str(short), 1, 3, bw)                       # 16000 samples < one 32768-sample block
    with Plan(PlanConfig(nchan=32, bw_mhz=[-bw], tscrunch=32)) as pl:
        for src in (empty, short):
            out = tmp_path / (src.name + ".fil")
            r = pl.run_scan([str(src)], str(out), source_name="X")
            h, off = sigproc.read_header(open(out, "rb").read())
            assert r["rows"] == 0 and os.path.getsize(out) == off and h.nchans == 32 and h.source_name == "X"
        r = pl.run_scan([str(short)], str(tmp_path / "late.fil"), start_s=5.0)
        assert r["rows"] == 0 and r["frames_per_if"] == 0
    nif, nfr = 32, 64
    vd, paths = [], []
    for i in range(nif):
        p = tmp_path / f"w_IF{i + 1}.vdif"
        vd.append(_write_vdif(str(p), nfr, 900 + i, bw, tone_frac=0.03 * i))
        paths.append(str(p))
    bws = [bw if (i + 1) % 2 == 0 else -bw for i in range(nif)]
    freqs = [1000.0 + i * bw for i in range(nif)]
    with Plan(PlanConfig(nchan=8, bw_mhz=bws, freq_mhz=freqs, tscrunch=64, rescale_interval_s=0.01)) as pl:
        r = pl.run_scan(paths, str(tmp_path / "wide.fil"))
    h, d = sigproc.read_fil(str(tmp_path / "wide.fil"))
    ref = o.base2fil({i + 1: vd[i] for i in range(nif)}, nif=nif, freq_lsb0=1000.0, bw=bw, nchan=8, tscrunch_factor=64,
                     rescale_interval_s=0.01)
    assert h.nchans == 256 and d.shape[0] == ref["data"].shape[0] == r["rows"] > 0
    assert np.abs(d[:, 0, :].astype(int) - ref["data"].astype(int)).max() <= 1


def test_c_program_through_the_abi(gpu, tmp_path):
    """examples/b2f_scan.c (plain C against include/b2f.h) writes the file the Python host writes."""
    import shutil
    import subprocess
    from frb_baseband_b200 import _lib
    from frb_baseband_b200.plan import Plan, PlanConfig
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe, libdir = str(tmp_path / "b2f_scan"), os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "b2f_scan.c"),
                           "-o", exe, "-L", libdir, "-lb2f", "-Wl,-rpath," + libdir])
    bw, paths = 32.0, []
    for i in range(2):
        p = tmp_path / f"c_IF{i + 1}.vdif"
        _write_vdif(str(p), 1300, 300 + i, bw, tone_frac=0.2 + 0.4 * i)
        paths.append(str(p))
    out_c, out_py = tmp_path / "c.fil", tmp_path / "py.fil"
    r = subprocess.run([exe, "--nchan", "128", "--tscrunch", "16", "--bw", "32", "--freq-lsb0", "1254", "--nsec", "0.3",
                        "--source", "R3", str(out_c)] + paths, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "2 IFs x 0.300 s" in r.stderr
    with Plan(PlanConfig(nchan=128, bw_mhz=[-bw, bw], freq_mhz=[1254.0, 1286.0], tscrunch=16)) as pl:
        pl.run_scan(paths, str(out_py), nsec=0.3, source_name="R3")
    assert open(out_c, "rb").read() == open(out_py, "rb").read()


def test_mode_a_concurrent_processes_share_the_gpu(gpu, tmp_path):
    """Mode A of INTEGRATION.md as base2fil.sh:60-66 runs it: one `process_vdif` OS process per IF, all at once on
    the same GPU, each into the FIFO base2fil made, with a splice-like reader joining them highest frequency first."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "bin", "process_vdif")
    nif, bw, nfr = 4, 16.0, 2600
    vd, procs, fifos = {}, [], []
    for i in range(1, nif + 1):
        p = tmp_path / f"ma_o8_no0003_IF{i}.vdif"
        vd[i] = _write_vdif(str(p), nfr, 600 + i, bw, tone_frac=0.1 * i)
        fifo = tmp_path / f"ma_o8_no0003_IF{i}.vdif_pol2.fil"
        os.mkfifo(fifo)                                                       # base2fil.sh:348-349
        fifos.append(str(fifo))
        freq = 1300.0 + (i - 1) * bw
        # the target of frb.conf carries --ra/--dec (frb.conf:3-6); without them the wrapper asks psrcat like the reference
        procs.append(subprocess.Popen([sys.executable, exe, "B0329+54", "--ra", "03:32:59.4", "--dec", "+54:34:43.3", str(p),
                                       "-f", str(freq), "-b", str(bw),
                                       "-l" if i % 2 else "-u", "--nchan", "32", "--nsec", "1", "--start", "0", "--force",
                                       "-t", "onsala85", "--pol", "2", "--tscrunch", "32", "--fil_out_dir", str(tmp_path),
                                       "--nbit=8"], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
    got = {}

    def reader(k, path):
        with open(path, "rb") as f:
            got[k] = f.read()

    threads = [threading.Thread(target=reader, args=(k, f), daemon=True) for k, f in enumerate(fifos)]
    for t in threads:
        t.start()
    for pr in procs:
        _, err = pr.communicate(timeout=120)
        assert pr.returncode == 0, err.decode()[-600:]
    for t in threads:
        t.join(60)
        assert not t.is_alive()
    ref = o.base2fil(vd, nif=nif, freq_lsb0=1300.0, bw=bw, nchan=32, tscrunch_factor=32, nsec=1.0)
    tiles = []
    for k in reversed(range(nif)):                                            # splice order: highest frequency first
        h, off = sigproc.read_header(got[k])
        tiles.append(np.frombuffer(got[k], np.uint8, offset=off).reshape(-1, 32))
        assert all(os.path.exists(f) for f in fifos)                          # FIFOs are never unlinked (:146-149)
    rows = np.concatenate(tiles, axis=1)
    assert rows.shape == ref["data"].shape
    assert np.abs(rows.astype(int) - ref["data"].astype(int)).max() <= 1


def test_peer_splice_equals_nccl_gather_on_two_gpus(gpu):
    """the one exchange of the path is the frequency splice (base2fil.sh:422).  Across GPUs it is done by peer
    stores over NVLink from the requantise kernel (PeerSplice + b2f_pull_strided); the rows must equal an NCCL gather of
    every rank's tile laid out highest frequency first.  Needs two GPUs on the box; skipped (not failed) on one."""
    import subprocess
    import sys
    from frb_baseband_b200 import _lib
    if _lib.lib().b2f_device_count() < 2:
        pytest.skip("one GPU on this box: the peer-store splice needs two")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29641", os.path.join(root, "tools", "check_peer_splice.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "equal=True" in r.stdout
