"""Build libpbg_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR.parent / "csrc"
INCLUDE = PKG_DIR.parent.parent / "include"
LIB_PATH = PKG_DIR / "libpbg_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
    "-Xptxas", "-v",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [INCLUDE / "pbg.h"]


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(s.stat().st_mtime > t for s in _sources())


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    return None


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libpbg_b200.so (there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, f"-I{INCLUDE}", "-o", str(LIB_PATH), str(CSRC / "pbg.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libpbg_b200.so")
    (PKG_DIR / "build_ptxas.log").write_text(proc.stdout + proc.stderr)
    return LIB_PATH


def build_variant(tag: str, defines: list[str]) -> Path:
    """A/B builds for same-box comparisons: libpbg_b200.<tag>.so with extra -D flags; select with PBG_LIB_PATH."""
    nvcc = find_nvcc()
    out = PKG_DIR / f"libpbg_b200.{tag}.so"
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], f"-I{INCLUDE}", "-o", str(out), str(CSRC / "pbg.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed building {out.name}")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:   # python -m pbg.build --variant TAG DEFINE[=V] ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
