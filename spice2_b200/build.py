"""Build the CUDA library in-tree: spice2_b200/libspice_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  Flags that matter for parity:
  * builtin_models.cu (everything that contains user functors) is compiled with -fmad=false so
    the functors' float expressions are evaluated operation by operation, as the IEEE-strict
    build of the reference does (DESIGN.md §parity contract);
  * host code is compiled with -ffp-contract=off (kahan_sum, fixed_probability constants).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
ROOT = HERE.parent
OUT = HERE / "libspice_b200.so"
OBJ = HERE / "build"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-std=c++20", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
          "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", f"-I{CSRC / 'include'}", f"-I{ROOT / 'include'}",
          f"-I{CSRC}"]
UNITS = [("runtime.cu", []), ("deliver.cu", ["-Xptxas", "-v"]), ("generator.cu", []), ("builtin_models.cu", ["-fmad=false"]), ("jump.cpp", [])]


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.rglob("*.cu")) + list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.h")) +
                    list(CSRC.rglob("*.cpp")) + [ROOT / "include" / "spice_b200.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = OBJ / "stamp"
    digest = _digest()
    if not force and OUT.exists() and stamp.exists() and stamp.read_text() == digest:
        return OUT
    if not force and OUT.exists() and os.environ.get("SPICE_PREBUILT") == "1":
        return OUT  # the library that travelled with the snapshot, whatever the sources say (gpurun calls)
    if not Path(NVCC).exists():
        if OUT.exists():
            return OUT  # GPU box without a toolkit: use the prebuilt library that travelled with the snapshot
        raise RuntimeError(f"nvcc not found at {NVCC} and no prebuilt {OUT.name}")
    OBJ.mkdir(exist_ok=True)
    # several ranks of one box may find the library stale at the same moment (torchrun): one of them builds, into private
    # names, and renames the result into place; the others wait for the lock and find the stamp up to date
    import fcntl

    with open(OBJ / "lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and OUT.exists() and stamp.exists() and stamp.read_text() == digest:
            return OUT
        return _build_locked(digest, stamp, verbose)


def _build_locked(digest: str, stamp: Path, verbose: bool) -> Path:
    tag = f".{os.getpid()}"
    procs = []
    for src, extra in UNITS:
        obj = OBJ / (src.rsplit(".", 1)[0] + tag + ".o")
        cmd = [NVCC, *COMMON, *extra, "-x", "cu", "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
        objs.append(str(obj))
    tmp_out = OUT.with_name(OUT.name + tag)
    cmd = [NVCC, "-shared", "-o", str(tmp_out), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for o in objs:
        Path(o).unlink(missing_ok=True)
    if r.returncode != 0:
        tmp_out.unlink(missing_ok=True)
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp_out, OUT)  # atomic: a process that has the old library mapped keeps it
    stamp.write_text(digest)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
