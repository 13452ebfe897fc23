"""Build libfluxb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libfluxb200.so")
SOURCES = ["api.cu", "bvh_build.cu", "render.cu", "render_regen.cu", "render_wave2.cu", "samplegen.cu"]

# -fmad=false: the Rust reference never contracts a*b+c; bit parity of hit distances and
# radiance depends on it.  Host code: -ffp-contract=off for the same reason.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-diag-suppress", "186",   # "pointless comparison of unsigned integer with zero": loops over counts that an instantiation fixes at 0
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2", "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "fluxb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra=(), out: str = LIB) -> str:
    """Compile libfluxb200.so.  `extra` adds nvcc flags and `out` another output path: A/B variants of
    compile-time kernel knobs (tools/build_variants.sh), selected at run time with FLUXB200_LIB."""
    if out == LIB and not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(out), exist_ok=True)
    ccbin = ["-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
    # one nvcc per translation unit, side by side (what a single nvcc call over all six does one after the other: a
    # minute), then the link
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    with tempfile.TemporaryDirectory(prefix="fluxb200_obj_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, src[:-3] + ".o")
            cmd = [nvcc_path(), *[f for f in NVCC_FLAGS if f != "-shared"], *extra, *ccbin, "-c", "-o", obj, os.path.join(CSRC, src)]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
            return obj
        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
            objs = list(pool.map(compile_one, SOURCES))
        cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", *ccbin, "-o", out, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "-f"]
    out = LIB
    if "-o" in args:
        k = args.index("-o")
        out = os.path.abspath(args[k + 1])
        del args[k:k + 2]
    print(build(force=True, verbose=True, extra=args, out=out))
