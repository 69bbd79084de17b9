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
