"""In-tree native build of libphyloacc_b200.so and the PhyloAcc-ST-compatible binary (nvcc, sm_100a only)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("PA_LIB_OUT", os.path.join(HERE, "libphyloacc_b200.so"))  # PA_LIB_OUT: build an A/B variant next to the product library
BIN = os.path.join(HERE, "PhyloAcc-ST-b200")
BIN_GBGC = os.path.join(HERE, "PhyloAcc-GBGC-b200")  # drop-in for the PhyloAcc-ST-GBGC binary (same source, -DPA_MAIN_GBGC)
LIB_SOURCES = ["pa_host.cpp", "pa_eval.cpp", "pa_app.cpp", "pa_capi.cpp", "pa_engine.cu", "pa_chain.cu", "pa_summary.cu", "pa_gbgc.cu", "pa_gbgc_chain.cu", "pa_gbgc_host.cpp", "pa_gbgc_app.cpp", "pa_gt.cu"]
NVCC = os.environ.get("PA_NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOSTCXX = "/usr/bin/g++"


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "phyloacc_b200.h"))
    objdir = os.path.join(HERE, "_build" + os.environ.get("PA_BUILD_TAG", ""))
    os.makedirs(objdir, exist_ok=True)
    if not (force is False and _newer(LIB, deps)):
        objs = []
        procs = []
        for s in srcs:
            o = os.path.join(objdir, os.path.basename(s) + ".o")
            objs.append(o)
            if not force and _newer(o, deps):
                continue
            cmd = [NVCC, "-ccbin", HOSTCXX, "-std=c++17", "-O3", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-fopenmp,-Wall",
                   "-x", "cu", "-dc" if False else "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            for extra in os.environ.get("PA_NVCC_FLAGS", "").split():
                cmd.insert(1, extra)
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for s, p in procs:
            out, _ = p.communicate()
            if verbose or p.returncode != 0:
                sys.stderr.write(out)
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
        cmd = [NVCC, "-ccbin", HOSTCXX, "-shared", *ARCH, "-Xcompiler", "-fPIC,-fopenmp", "-o", LIB, *objs, "-lgomp"]
        subprocess.run(cmd, check=True)
    main_src = os.path.join(CSRC, "pa_main.cpp")
    if os.path.exists(main_src) and not _newer(BIN, [main_src, LIB]):
        # the binary dlopen()s the library only after argument checking (zero-argument probe must not touch CUDA)
        subprocess.run([HOSTCXX, "-O2", "-std=c++17", main_src, "-o", BIN, "-ldl"], check=True)
    if os.path.exists(main_src) and not _newer(BIN_GBGC, [main_src, LIB]):
        subprocess.run([HOSTCXX, "-O2", "-std=c++17", "-DPA_MAIN_GBGC", main_src, "-o", BIN_GBGC, "-ldl"], check=True)
    return LIB


if __name__ == "__main__":
    build_native(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
