"""
Builds libxcb200.so (hand-written CUDA for sm_100a + the C ABI of
include/xcb200.h) in-tree with nvcc.  No GPU is needed to build.

    python -m xcontour_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libxcb200.so")
SOURCES = ["api.cu", "levels.cu", "hist.cu", "bin_rows.cu", "contour_ops.cu", "lwa.cu", "lwa_fx.cu", "lwa_cols.cu", "grad2.cu", "epilogue.cu", "equal_area.cu", "fused.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "xcb200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: build an experimental libxcb200_<variant>.so with extra -D
    flags (selected at run time with XCB200_LIB=<path>); used for A/B timing only."""
    lib = LIB if variant is None else os.path.join(HERE, "libxcb200_%s.so" % variant)
    if variant is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    suffix = ".o" if variant is None else ".%s.o" % variant
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", suffix))
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (src, out))
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
    with open(os.path.join(CSRC, "build.log" if variant is None else "build.%s.log" % variant), "w") as f:
        f.write("\n".join(log))
    if failed:
        raise RuntimeError("libxcb200 build failed")
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
