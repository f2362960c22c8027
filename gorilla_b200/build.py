"""Build libgorilla_b200.so (CUDA kernels for sm_100a + host mesh builder) in-tree.

    python -m gorilla_b200.build [--force]

Translation units are compiled in parallel (nvcc cross-compiles without a GPU) and linked into
gorilla_b200/lib/libgorilla_b200.so.  --fmad=false is part of the contract: the reference ISA has no
FMA and the visited-tetra sequence is only reproducible with separately rounded multiplies and adds.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
# GORILLA_VARIANT=name builds lib/libgorilla_b200_name.so from its own object directory (tuning experiments with
# GORILLA_NVCC_EXTRA; select it at run time with GORILLA_B200_LIB)
_VARIANT = os.environ.get("GORILLA_VARIANT", "")
OUT_DIR = ROOT / "lib"
OBJ_DIR = ROOT / "lib" / ("obj_" + _VARIANT if _VARIANT else "obj")
LIB = OUT_DIR / ("libgorilla_b200_" + _VARIANT + ".so" if _VARIANT else "libgorilla_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off", "-Xptxas", "-v",
]
_EXTRA = os.environ.get("GORILLA_NVCC_EXTRA", "").split()
if any(f.startswith("--fmad") for f in _EXTRA):   # a variant that allows contraction (measurement only: not bit-exact)
    NVCC_FLAGS = [f for f in NVCC_FLAGS if not f.startswith("--fmad")]
NVCC_FLAGS += _EXTRA
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas"]

# longest compiles first (order 4 takes ~2 min per unit)
CU_SOURCES = ["gb_orbit_k4ax.cu", "gb_orbit_k3ax.cu", "gb_orbit_k4x.cu", "gb_orbit_k4a.cu", "gb_orbit_k4t.cu", "gb_orbit_k4.cu", "gb_orbit_k4p.cu", "gb_orbit_k3p.cu", "gb_orbit_k2p.cu", "gb_orbit_k3x.cu", "gb_orbit_k3a.cu",
              "gb_orbit_k3t.cu", "gb_orbit_k3.cu", "gorilla_b200.cu", "gb_diag.cu", "gb_orbit_rk.cu", "gb_orbit_rkx.cu", "gb_orbit_k2ax.cu", "gb_orbit_k1ax.cu", "gb_orbit_k2x.cu", "gb_orbit_k2a.cu",
              "gb_orbit_k2t.cu", "gb_orbit_k2.cu", "gb_orbit_k1x.cu", "gb_orbit_k1a.cu", "gb_orbit_k1t.cu", "gb_orbit_k1.cu"]
CPP_SOURCES = ["host/mesh_api.cpp", "host/mesh_common.cpp", "host/mesh_analytic.cpp", "host/mesh_vmec.cpp", "host/mesh_efit.cpp", "host/mesh_soledge3x.cpp", "host/mesh_efit_flux.cpp"]


def _headers_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.hpp")) + list(CSRC.rglob("*.h"))
                    + [ROOT.parent / "include" / "gorilla_b200.h"]):
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, hdr_digest: str, force: bool) -> tuple[Path, str]:
    path = CSRC / src
    obj = OBJ_DIR / (src.replace("/", "_") + ".o")
    stamp = obj.with_suffix(".stamp")
    digest = hashlib.sha256(path.read_bytes() + hdr_digest.encode()).hexdigest()
    if not force and obj.exists() and stamp.exists() and stamp.read_text().split("\n")[0] == digest:
        return obj, None
    if src.endswith(".cu"):
        cmd = [NVCC, *NVCC_FLAGS, "-c", str(path), "-o", str(obj)]
    else:
        cmd = [HOST_CXX, *CXX_FLAGS, "-c", str(path), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"compile failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest + "\n" + r.stderr)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    hd = _headers_digest()
    srcs = CU_SOURCES + CPP_SOURCES
    # GORILLA_VARIANT_UNITS=a.cu,b.cu (with GORILLA_VARIANT): only these units are compiled with the variant's flags, every
    # other object is taken from the main build (tuning experiments on one kernel need not recompile all twenty units)
    only = [u for u in os.environ.get("GORILLA_VARIANT_UNITS", "").split(",") if u]
    borrowed = []
    if _VARIANT and only:
        main_obj = ROOT / "lib" / "obj"
        borrowed = [str(main_obj / (s.replace("/", "_") + ".o")) for s in srcs if s not in only]
        missing = [b for b in borrowed if not Path(b).exists()]
        if missing:
            raise RuntimeError(f"GORILLA_VARIANT_UNITS needs the main build first (missing {missing[0]})")
        srcs = [s for s in srcs if s in only]
    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, hd, force), srcs))
    objs = [str(o) for o, _ in results] + borrowed
    rebuilt = any(log is not None for _, log in results) or not LIB.exists()  # None = object was up to date
    if verbose:
        for _, log in results:
            if log:
                sys.stderr.write(log)
    if rebuilt or force:
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOST_CXX,
               "-Xcompiler", "-fPIC,-fopenmp", "-o", str(LIB), *objs, "-lgomp", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return LIB


def ptxas_report() -> str:
    """Register / spill report of the last compile of every .cu (from the stamp files)."""
    out = []
    for src in CU_SOURCES:
        stamp = OBJ_DIR / (src.replace("/", "_") + ".stamp")
        if stamp.exists():
            out.append(f"== {src}\n" + "\n".join(stamp.read_text().split("\n")[1:]))
    return "\n".join(out)


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
