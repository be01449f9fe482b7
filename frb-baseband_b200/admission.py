"""Admission control for compatibility mode A (one `process_vdif` per IF, all sharing one GPU).

The reference throttles by counting processes *named* `digifil` (`pwait` in /root/reference/base2fil.sh:21-28, used at
:404-405; /root/reference/online-deamon.sh:50 does the same for whole scans).  A replacement that is not called digifil
makes that count zero, so nothing would stop several scans' worth of subband processes from piling onto one GPU.  The
same rule is restated here on lock files: at most B2F_MAX_CONCURRENT holders per device; a process that finds every
slot taken waits, like `pwait` does.  Unset or 0 = no limit (mode B, one call per scan, needs none)."""
from __future__ import annotations

import fcntl
import os
import time


class GpuSlot:
    def __init__(self, device: int = 0, limit: int | None = None, lock_dir: str | None = None, poll_s: float = 0.1):
        self.limit = int(os.environ.get("B2F_MAX_CONCURRENT", "0")) if limit is None else int(limit)
        self.dir = lock_dir or os.environ.get("B2F_LOCK_DIR", "/tmp")
        self.device, self.poll_s, self.fd, self.slot = device, poll_s, None, None

    def acquire(self, timeout_s: float | None = None) -> int | None:
        if self.limit <= 0:
            return None
        t0 = time.monotonic()
        while True:
            for k in range(self.limit):
                fd = os.open(os.path.join(self.dir, f"b2f_gpu{self.device}_slot{k}.lock"), os.O_CREAT | os.O_RDWR, 0o666)
                try:
                    fcntl.flock(fd, fcntl.LOCK_EX | fcntl.LOCK_NB)
                except OSError:
                    os.close(fd)
                    continue
                self.fd, self.slot = fd, k
                return k
            if timeout_s is not None and time.monotonic() - t0 > timeout_s:
                raise TimeoutError(f"no free b2f slot on GPU {self.device} after {timeout_s} s (B2F_MAX_CONCURRENT={self.limit})")
            time.sleep(self.poll_s)

    def release(self):
        if self.fd is not None:
            os.close(self.fd)            # closing drops the flock, also when the process dies
            self.fd = self.slot = None

    def __enter__(self):
        self.acquire()
        return self

    def __exit__(self, *a):
        self.release()
