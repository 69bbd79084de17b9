"""In-tree build driver: `make` at the repo root -> bagua_net_b200/lib/*.so."""
from __future__ import annotations

import os
import subprocess
import sys

from . import LIB_DIR, LIB_NAME, REPO_ROOT


def _newest_source_mtime() -> float:
    newest = 0.0
    for root in ("csrc", "include"):
        for dp, _, files in os.walk(os.path.join(REPO_ROOT, root)):
            for f in files:
                if f.endswith((".cc", ".cu", ".h", ".cuh")):
                    newest = max(newest, os.path.getmtime(os.path.join(dp, f)))
    mk = os.path.join(REPO_ROOT, "Makefile")
    if os.path.exists(mk):
        newest = max(newest, os.path.getmtime(mk))
    return newest


def is_stale() -> bool:
    so = os.path.join(LIB_DIR, LIB_NAME)
    return not os.path.exists(so) or os.path.getmtime(so) < _newest_source_mtime()


def build_native(verbose: bool = False, jobs: int | None = None, force: bool = False) -> str:
    so = os.path.join(LIB_DIR, LIB_NAME)
    if not force and not is_stale():
        return so
    # One builder at a time across PROCESSES: under torchrun every rank of a fresh checkout gets here at once, and
    # `make` in one build directory from several processes corrupts objects.  The others wait on the lock, then find
    # the library up to date.  (make links to a temporary name and renames, so nobody dlopens a half-written file.)
    import fcntl

    os.makedirs(os.path.join(REPO_ROOT, "build"), exist_ok=True)
    with open(os.path.join(REPO_ROOT, "build", ".build.lock"), "w") as lockf:
        fcntl.flock(lockf, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return so
            return _run_make(so, verbose, jobs)
        finally:
            fcntl.flock(lockf, fcntl.LOCK_UN)


def _run_make(so: str, verbose: bool, jobs: int | None) -> str:
    jobs = jobs or min(16, os.cpu_count() or 4)
    cmd = ["make", "-C", REPO_ROOT, f"-j{jobs}"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout[-8000:])
    if proc.returncode != 0:
        raise RuntimeError(f"native build failed (exit {proc.returncode}); see output above")
    if not os.path.exists(so):
        raise RuntimeError(f"native build finished but {so} is missing")
    return so
