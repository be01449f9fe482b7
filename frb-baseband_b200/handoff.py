"""What happens to a spliced filterbank after it is written.

Mirrors of the reference's downstream hand-off, as data (argv lists / message strings) so that the caller
decides whether the external tools exist on this host:

  * FETCH submission: `pika_send.py -q stage01_queue -m "<fil> <flag>"` (/root/reference/base2fil.sh:118-122,
    issued after a successful splice `:420-435`);
  * scan clean-up: split VDIF files removed unless keepVDIF (`:423-425,440-442`);
  * pulsar fold + plots when psrcat knows the target (`:452-497`).

`base2fil.run_scan` calls `after_scan` with the same ordering as the shell: the message goes out only after the
output file is complete, the VDIF files go away only after a successful run.
"""
from __future__ import annotations

import glob
import os
import shlex
import subprocess
import sys
from typing import Callable

FETCH_PYTHON = "/home/franz/.conda/envs/fetch/bin/python"               # base2fil.sh:121
FETCH_SENDER = "/home/franz/software/src/greenburst/pika_send.py"      # base2fil.sh:121
FETCH_QUEUE = "stage01_queue"


def fetch_message(fil_path: str, flag_file: str = "") -> str:
    """Body of the queue message: '<filterbank> <flag file>'; the flag file may be empty (base2fil.sh:119-121)."""
    return f"{fil_path} {flag_file}"


def fetch_argv(fil_path: str, flag_file: str = "", *, python: str = FETCH_PYTHON, sender: str = FETCH_SENDER,
               queue: str = FETCH_QUEUE) -> list[str]:
    return [python, sender, "-q", queue, "-m", fetch_message(fil_path, flag_file)]


def target_name(target: str) -> str:
    """`frb.conf` targets may carry `--ra .. --dec ..`; the fold step keeps what precedes the first '-' and drops
    trailing blanks (base2fil.sh:452-454)."""
    return target.split("-")[0].rstrip(" ")


def split_vdif_globs(experiment: str, st: str, scanname: str, workdir_odd: str, workdir_even: str) -> list[str]:
    """patterns the reference removes after the splice (base2fil.sh:424-425)"""
    tail = f"{experiment}_{st}_no0{scanname}_IF*.vdif"
    return [os.path.join(workdir_even, tail), os.path.join(workdir_odd, tail)]


def fold_commands(target: str, station: str, fil_path: str, pol: int) -> list[str]:
    """Shell lines of the fold-and-plot step for one scan (base2fil.sh:474-493).  The caller first runs
    `psrcat -e <target> > <target>.psrcat.par` and only continues when that succeeds (`:463-464`); BSGR is
    skipped by name (`:458-460`)."""
    t = target_name(target)
    if t == "BSGR":
        return []
    cmds = [
        f"dspsr -E {t}.psrcat.par -L 10 -A -k {station} -d1 {fil_path} -O {fil_path} -t 8",
        f"psrplot -pF -D /CPS -c x:unit=s {fil_path}.ar -j dedisperse,tscrunch,pscrunch,\"fscrunch 128\"",
        f"mv pgplot.ps {fil_path}.ps",
    ]
    if pol == 4:
        cmds += [
            f"psrplot -N 2x2 -D /CPS {fil_path}.ar -j tscrunch,dedisperse,\"fscrunch 128\" "
            "-p freq+ -c ':0:pol=0' -p freq+ -c ':1:pol=1' -p freq+ -c ':2:x:unit=ms' -j :2:pscrunch "
            "-p Scyl -j :3:fscrunch",
            f"mv pgplot.ps {fil_path}_fullPol.ps",
        ]
    return cmds


def after_scan(fil_path: str, *, flag_file: str = "", submit2fetch: bool = False, keep_vdif: bool = True,
               vdif_globs: list[str] | None = None, send: Callable[[list[str]], int] | None = None,
               log=None) -> dict:
    """The `&&` chain behind `splice ... > out` (base2fil.sh:420-447): remove the split VDIF files unless kept,
    then submit to FETCH.  `send` runs the sender argv (default: subprocess.call); returns what was done."""
    log = log or (lambda m: print(m, file=sys.stderr))
    done = {"removed": [], "submitted": None}
    if not os.path.exists(fil_path):
        raise FileNotFoundError(fil_path)
    if not keep_vdif:
        for pat in vdif_globs or []:
            for f in sorted(glob.glob(pat)):
                os.remove(f)
                done["removed"].append(f)
    if submit2fetch:
        argv = fetch_argv(fil_path, flag_file)
        rc = (send or subprocess.call)(argv)
        if rc != 0:
            raise RuntimeError("FETCH submission failed: " + " ".join(shlex.quote(a) for a in argv))
        done["submitted"] = fetch_message(fil_path, flag_file)
        log(f"Submitted {fil_path} {flag_file} to fetch")
    return done
