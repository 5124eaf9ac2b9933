"""In-tree build of libgq_b200.so (nvcc, sm_100a only).

    python vq-vae-from-gaussian-vae_b200/build_ext.py [--force] [-v]

Stand-alone on purpose (it must not import gq_b200, whose import loads the library).  The
shared object lands in gq_b200/ so that it travels with the source tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent / "gq_b200"
CSRC = PKG_DIR.parent / "csrc"
LIB = PKG_DIR / "libgq_b200.so"
SOURCES = ["gq_api.cu", "gq_kernels.cu", "gq_kl.cu", "gq_tc.cu", "gq_tokens.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libgq_b200.so cannot be built")
    return exe


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(PKG_DIR.parent.parent / "include" / "gq_b200.h")
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu to an object file (in parallel: gq_tc.cu alone is most of the time) and link the shared
    library.  No relocatable device code is involved, so plain objects link."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = CSRC / "_obj"
    obj_dir.mkdir(exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src: str):
        obj = obj_dir / (Path(src).stem + ".o")
        cmd = [_nvcc(), *compile_flags, "-c", "-o", str(obj), str(CSRC / src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        return src, obj, subprocess.run(cmd, cwd=str(CSRC), capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    failed = False
    for src, _, res in results:
        if verbose or res.returncode:
            sys.stderr.write(res.stdout + res.stderr)
        failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", str(LIB),
            *[str(o) for _, o, _ in results]]
    res = subprocess.run(link, cwd=str(CSRC), capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError(f"nvcc link failed with exit code {res.returncode}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
