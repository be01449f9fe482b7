"""Scan-level driver: every subband of a scan through one GPU plan, spliced in device memory.

This replaces lines 387-448 of /root/reference/base2fil.sh -- the fan-out of one background
`process_vdif`/`digifil` per IF into FIFOs plus the `splice` that joins them -- with a single
call that writes the final file base2fil names
    ${outdir}/${experiment}_${st}_no0${scanname}_IFall_vdif_pol${pol}.fil        (base2fil.sh:389)
from the split files
    ${workdir_odd|even}/${experiment}_${st}_no0${scanname}_IF${i}.vdif          (base2fil.sh:336,353)
Everything around it in base2fil.sh (jive5ab split, FETCH submission, folding) is untouched;
INTEGRATION.md shows the ten-line change that calls this instead of run_process_vdif+splice.
"""
from __future__ import annotations

import os
import sys

from . import handoff, sigproc, vdif
from .conf import FrbConf, read_conf
from .plan import Plan, PlanConfig, pol_mode_from_reference

#: /root/reference/base2fil.sh:149-169
STATION_CODES = dict(zip(
    "onsala85 onsala60 srt wsrt effelsberg torun tianma irbene irbene16 medicina noto urumqi badary svetloe".split(),
    "o8 o6 sr wb ef tr t6 ir ib mc nt ur bd sv".split()))


def station_code(station: str) -> str:
    try:
        return STATION_CODES[station.lower()]
    except KeyError:
        raise ValueError(f"Station {station.lower()} not known. Options are: ({' '.join(STATION_CODES)})")


def scan_files(cfg: FrbConf, scanname: str, workdir_odd: str | None = None, workdir_even: str | None = None):
    """{IF number: split VDIF path}; odd IFs live under workdir_odd, even under workdir_even."""
    st = station_code(cfg.station)
    odd = workdir_odd or os.path.join(os.path.expandvars(cfg.workdir_odd_base), cfg.experiment)
    even = workdir_even or os.path.join(os.path.expandvars(cfg.workdir_even_base), cfg.experiment)
    name = "%03d" % int(scanname)
    return {i: os.path.join(odd if i % 2 else even, f"{cfg.experiment}_{st}_no0{name}_IF{i}.vdif")
            for i in range(1, int(cfg.nif) + 1)}


def spliced_name(cfg: FrbConf, scanname: str) -> str:
    return f"{cfg.experiment}_{station_code(cfg.station)}_no0{'%03d' % int(scanname)}_IFall_vdif_pol{cfg.pol}.fil"


def seconds_in_file(path: str, info: vdif.FrameInfo, datarate_mbps: int, nif: int) -> int:
    """nsec exactly as base2fil.sh:395-402 derives it with integer `bc` arithmetic"""
    fps_all = datarate_mbps * 1000000 // 8 // info.payload_bytes
    return vdif.nsec_in_file(os.path.getsize(path), info.frame_bytes, fps_all // nif)


def run_scan(files: dict[int, str], out_path: str, *, bw: float, freq_lsb0: float, nchan: int, tscrunch: int = 1,
             pol: int = 2, nbit: int = 8, start: float = 0.0, nsec: float | None = None, keep_bandpass: bool = False,
             source: str = "unknown", ra: str | None = None, dec: str | None = None, telescope: str = "",
             device: int = 0, chunk_units: int = 0, verbose: bool = True, dm: float = 0.0, coherent: bool = False) -> dict:
    """All IFs of one scan -> one band-ordered SIGPROC filterbank.  Returns counters + geometry."""
    ifs = sorted(files)
    nif = len(ifs)
    freqs = [freq_lsb0 + (i - 1) * bw for i in ifs]                    # base2fil.sh:54,65,254
    bws = [bw if i % 2 == 0 else -bw for i in ifs]                     # odd = LSB (-l), even = USB (-u)
    infos = []
    for i in ifs:
        with open(files[i], "rb") as f:
            infos.append(vdif.parse_header(f.read(32)))
    info = infos[0]
    for k, other in enumerate(infos):
        if (other.frame_bytes, other.nbit, other.header_bytes) != (info.frame_bytes, info.nbit, info.header_bytes):
            raise ValueError(f"{files[ifs[k]]}: frame geometry differs from {files[ifs[0]]}")
    cfg = PlanConfig(nchan=nchan, bw_mhz=bws, freq_mhz=freqs, tscrunch=max(1, tscrunch),
                     pol_mode=pol_mode_from_reference(pol), out_nbit=nbit, in_nbit=info.nbit,
                     frame_bytes=info.frame_bytes, header_bytes=info.header_bytes, keep_bandpass=keep_bandpass,
                     device=device, chunk_units=chunk_units, dm=dm, coherent=coherent and dm > 0)
    with Plan(cfg) as pl:
        r = pl.run_scan([files[i] for i in ifs], out_path, start_s=start, nsec=nsec, source_name=source,
                        telescope_id=sigproc.TELESCOPE_IDS.get(telescope.lower(), 0),
                        src_raj=sigproc.sexagesimal_to_sigproc(ra), src_dej=sigproc.sexagesimal_to_sigproc(dec))
        tsamp = pl.tsamp_s
    return _report(r, f"{nif} IFs x {r['seconds_of_data']:.3f} s", out_path, nif * nchan, tsamp, verbose)


def _report(r: dict, what: str, out_path: str, nchans: int, tsamp: float, verbose: bool) -> dict:
    c = r["counters"]
    if verbose:
        print(f"b2f: {what} -> {out_path}: {r['rows']} samples x {nchans} channels in {r['wall_s']:.2f} s; "
              f"frames ok/invalid/fill/bad = {c['frames_ok']}/{c['frames_invalid']}/{c['frames_with_fill']}/"
              f"{c['frames_badhdr']}", file=sys.stderr)
    return {"rows": r["rows"], "counters": c, "nchans": nchans, "tsamp_s": tsamp, "wall_s": r["wall_s"]}


def base2fil(conf_path: str, *, device: int = 0, workdir_odd: str | None = None, workdir_even: str | None = None,
             outdir: str | None = None, send=None) -> list[str]:
    """Filterbank stage of `base2fil <conf>` for scans whose split VDIF files already exist, followed by the
    reference's per-scan hand-off (`submit2fetch`, `keepVDIF`, `flagFile`: base2fil.sh:420-447).  `send` replaces the
    FETCH sender process (see handoff.after_scan)."""
    cfg = read_conf(conf_path)
    outdir = outdir or os.path.join(os.path.expandvars(cfg.outdir_base), cfg.experiment)
    os.makedirs(outdir, exist_ok=True)
    targ = cfg.target_args()
    source = targ[0] if targ else "unknown"
    ra = dec = None
    for k, tok in enumerate(targ):
        if tok == "--ra" and k + 1 < len(targ):
            ra = targ[k + 1]
        elif tok.startswith("--ra="):
            ra = tok[5:]
        elif tok == "--dec" and k + 1 < len(targ):
            dec = targ[k + 1]
        elif tok.startswith("--dec="):
            dec = tok[6:]
    done = []
    for scanname in cfg.scannames:
        files = scan_files(cfg, scanname, workdir_odd, workdir_even)
        out_path = os.path.join(outdir, spliced_name(cfg, scanname))
        last = files[int(cfg.nif)]
        if os.path.getsize(last) == 0:                      # base2fil.sh:391-394
            open(out_path, "wb").close()
            continue
        with open(last, "rb") as f:
            info = vdif.parse_header(f.read(32))
        nsec = seconds_in_file(last, info, cfg.datarate, int(cfg.nif))
        run_scan(files, out_path, bw=float(cfg.bw), freq_lsb0=float(cfg.freqLSB_0), nchan=int(cfg.nchan),
                 tscrunch=int(cfg.tscrunch), pol=int(cfg.pol), nbit=int(cfg.nbit), start=float(cfg.start), nsec=nsec,
                 keep_bandpass=int(cfg.keepBP) > 0, source=source, ra=ra, dec=dec, telescope=cfg.station, device=device)
        odd = os.path.dirname(files[1])
        even = os.path.dirname(files[2]) if int(cfg.nif) > 1 else odd
        handoff.after_scan(out_path, flag_file=str(cfg.flagFile), submit2fetch=int(cfg.submit2fetch) != 0,
                           keep_vdif=int(cfg.keepVDIF) != 0, send=send,
                           vdif_globs=handoff.split_vdif_globs(cfg.experiment, station_code(cfg.station),
                                                               "%03d" % int(scanname), odd, even))
        done.append(out_path)
    return done


def run_scan_raw(raw_path: str, out_path: str, *, mode: str, nif: int, bw: float, freq_lsb0: float, nchan: int,
                 tscrunch: int = 1, pol: int = 2, nbit: int = 8, start: float = 0.0, nsec: float | None = None,
                 keep_bandpass: bool = False, flip_if: bool = False, source: str = "unknown", ra: str | None = None,
                 dec: str | None = None, telescope: str = "", device: int = 0, verbose: bool = True) -> dict:
    """Raw multi-BBC recording -> band-ordered filterbank in one pass: the corner turn that
    base2fil.sh:334-368 delegates to jive5ab (`spif2file`, recipes of spif2file.sh:31-98) is done on the
    GPU, so no split files are written.  `mode` is base2fil's mode string (base2fil.sh:308-318)."""
    from . import spif

    W, bits = spif.recipe_for_mode(mode, nif, flip_if)
    ifs = list(range(1, nif + 1))
    freqs = [freq_lsb0 + (i - 1) * bw for i in ifs]
    bws = [bw if i % 2 == 0 else -bw for i in ifs]
    frame_bytes, header_bytes, raw_format = spif.frame_geometry(mode)
    if not raw_format:                                        # VDIF says what it is (legacy headers are 16 bytes)
        with open(raw_path, "rb") as f:
            info = vdif.parse_header(f.read(32))
        frame_bytes, header_bytes = info.frame_bytes, info.header_bytes
    cfg = PlanConfig(nchan=nchan, bw_mhz=bws, freq_mhz=freqs, tscrunch=max(1, tscrunch),
                     pol_mode=pol_mode_from_reference(pol), out_nbit=nbit, in_nbit=len(bits[0]) // 2, frame_bytes=frame_bytes,
                     header_bytes=header_bytes, keep_bandpass=keep_bandpass, device=device,
                     raw_word_bits=W, raw_bits=bits, raw_format=raw_format)
    with Plan(cfg) as pl:
        r = pl.run_scan([raw_path], out_path, start_s=start, nsec=nsec, source_name=source,
                        telescope_id=sigproc.TELESCOPE_IDS.get(telescope.lower(), 0),
                        src_raj=sigproc.sexagesimal_to_sigproc(ra), src_dej=sigproc.sexagesimal_to_sigproc(dec))
        tsamp = pl.tsamp_s
    return _report(r, f"raw {mode}", out_path, nif * nchan, tsamp, verbose)


def raw_mode(cfg: FrbConf, raw_path: str | None = None) -> str:
    """The mode string base2fil hands to spif2file (base2fil.sh:308-318): `VDIF_<payload>-<Mbps>-<nbbc>-<nbits>` with the
    payload read from the recording's first header, or `MARK5B-<Mbps>-<nbbc>-<nbits>` when the config sets isMark5b."""
    from . import spif

    nbbc = 2 * int(cfg.nif)
    if int(cfg.isMark5b):
        return f"MARK5B-{cfg.datarate}-{nbbc}-{int(cfg.nbits)}"
    with open(raw_path, "rb") as f:
        info = vdif.parse_header(f.read(32))
    return spif.mode_string(info.payload_bytes, cfg.datarate, nbbc, int(cfg.nbits))


def raw_window(cfg: FrbConf, k: int, raw_path: str, mode: str) -> tuple[float, float]:
    """(start_s, nsec) of scan k inside its raw recording.  spif2file.sh:144-151 skips to one frame before the first whole
    second (`fps - start_frame - 1` frames; start_frame is 0 for Mark5B) plus `skips[k]` seconds and splits `lengths[k]`
    seconds; process_vdif then runs digifil with -S `start` over what is left (base2fil.sh:61-64,395-402)."""
    from . import spif

    fb, hb, fmt = spif.frame_geometry(mode)
    W, _ = spif.recipe_for_mode(mode, int(cfg.nif))
    start_frame = 0
    if not fmt:
        with open(raw_path, "rb") as f:
            info = vdif.parse_header(f.read(32))
        fb, hb, start_frame = info.frame_bytes, info.header_bytes, info.frame_nr
    fps = round(2 * float(cfg.bw) * 1e6 / ((fb - hb) * 8 // W))
    lead = (fps - start_frame - 1) / fps
    start = float(cfg.start)
    return lead + float(cfg.skips[k]) + start, float(cfg.lengths[k]) - start


def base2fil_raw(conf_path: str, *, vbsdir: str | None = None, outdir: str | None = None, device: int = 0,
                 verbose: bool = True, send=None) -> list[str]:
    """Mode C from a frb.conf: for every scan, the raw recording
    `${vbsdir}/${experiment}_${st}_no0${scan}` (spif2file.sh:142) goes through the GPU corner turn straight into
    `${outdir}/${experiment}_${st}_no0${scanname}_IFall_vdif_pol${pol}.fil` -- no jive5ab split, no split files, no FIFOs.
    The mode (VDIF or Mark5B, word size, bits per sample), flipIF, the skip / length window and everything process_vdif is
    told come from the same keys base2fil.sh reads."""
    cfg = read_conf(conf_path)
    st = station_code(cfg.station)
    vbsdir = vbsdir or os.path.join(os.path.expandvars(cfg.vbsdir_base), cfg.experiment)
    outdir = outdir or os.path.join(os.path.expandvars(cfg.outdir_base), cfg.experiment)
    os.makedirs(outdir, exist_ok=True)
    targ = cfg.target_args()
    source = targ[0] if targ else "unknown"
    ra = dec = None
    for k, tok in enumerate(targ):
        if tok == "--ra" and k + 1 < len(targ):
            ra = targ[k + 1]
        elif tok.startswith("--ra="):
            ra = tok[5:]
        elif tok == "--dec" and k + 1 < len(targ):
            dec = targ[k + 1]
        elif tok.startswith("--dec="):
            dec = tok[6:]
    done = []
    for k, scanname in enumerate(cfg.scannames):
        raw_path = os.path.join(vbsdir, f"{cfg.experiment}_{st}_no0{cfg.scans[k]}")
        mode = raw_mode(cfg, raw_path)
        start_s, nsec = raw_window(cfg, k, raw_path, mode)
        out_path = os.path.join(outdir, spliced_name(cfg, scanname))
        run_scan_raw(raw_path, out_path, mode=mode, nif=int(cfg.nif), bw=float(cfg.bw), freq_lsb0=float(cfg.freqLSB_0),
                     nchan=int(cfg.nchan), tscrunch=int(cfg.tscrunch), pol=int(cfg.pol), nbit=int(cfg.nbit), start=start_s, nsec=nsec,
                     keep_bandpass=int(cfg.keepBP) > 0, flip_if=int(cfg.flipIF) != 0, source=source, ra=ra, dec=dec,
                     telescope=cfg.station, device=device, verbose=verbose)
        handoff.after_scan(out_path, flag_file=str(cfg.flagFile), submit2fetch=int(cfg.submit2fetch) != 0, keep_vdif=True, send=send,
                           vdif_globs=[])
        done.append(out_path)
    return done


if __name__ == "__main__":
    # base2fil <frb.conf>: split files -> filterbank (mode B); base2fil --raw <frb.conf>: raw recordings -> filterbank (mode C)
    _args = [a for a in sys.argv[1:] if a != "--raw"]
    for p in (base2fil_raw if "--raw" in sys.argv[1:] else base2fil)(_args[0]):
        print(p)
