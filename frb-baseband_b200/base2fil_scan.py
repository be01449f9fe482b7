"""`python -m frb_baseband_b200.base2fil_scan <frb.conf> <scanname> [--workdir-odd D --workdir-even D --outdir D]`

One scan's filterbank stage (what base2fil.sh:404-448 does with nif digifils + splice) as a single
command, for the ten-line base2fil.sh change shown in INTEGRATION.md section 3."""
import argparse
import os
import sys

from . import base2fil, vdif
from .conf import read_conf


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("conf")
    ap.add_argument("scanname")
    ap.add_argument("--workdir-odd")
    ap.add_argument("--workdir-even")
    ap.add_argument("--outdir")
    ap.add_argument("--device", type=int, default=int(os.environ.get("B2F_DEVICE", "0")))
    a = ap.parse_args(argv)
    cfg = read_conf(a.conf)
    outdir = a.outdir or os.path.join(os.path.expandvars(cfg.outdir_base), cfg.experiment)
    os.makedirs(outdir, exist_ok=True)
    files = base2fil.scan_files(cfg, a.scanname, a.workdir_odd, a.workdir_even)
    out_path = os.path.join(outdir, base2fil.spliced_name(cfg, a.scanname))
    last = files[int(cfg.nif)]
    if os.path.getsize(last) == 0:                       # base2fil.sh:391-394
        open(out_path, "wb").close()
        return 0
    with open(last, "rb") as f:
        info = vdif.parse_header(f.read(32))
    targ = cfg.target_args()
    ra = dec = None
    for k, tok in enumerate(targ):
        if tok in ("--ra", "--dec") and k + 1 < len(targ):
            ra, dec = (targ[k + 1], dec) if tok == "--ra" else (ra, targ[k + 1])
        elif tok.startswith("--ra="):
            ra = tok[5:]
        elif tok.startswith("--dec="):
            dec = tok[6:]
    base2fil.run_scan(files, out_path, bw=float(cfg.bw), freq_lsb0=float(cfg.freqLSB_0), nchan=int(cfg.nchan),
                      tscrunch=int(cfg.tscrunch), pol=int(cfg.pol), nbit=int(cfg.nbit), start=float(cfg.start),
                      nsec=base2fil.seconds_in_file(last, info, cfg.datarate, int(cfg.nif)),
                      keep_bandpass=int(cfg.keepBP) > 0, source=targ[0] if targ else "unknown", ra=ra, dec=dec,
                      telescope=cfg.station, device=a.device)
    print(out_path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
