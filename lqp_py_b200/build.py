"""Build the sm_100a CUDA library in-tree: ``python -m lqp_py_b200.build``.

Produces ``lqp_py_b200/_lqpb.so`` (git-ignored, travels to the GPU box with the snapshot).
nvcc cross-compiles without a GPU.  The library exports only the C ABI of ``include/lqpb.h``.
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "_lqpb.so")
STAMP = os.path.join(PKG, "_lqpb.stamp")
SOURCES = ["abi.cu", "scale.cu", "factor.cu", "tcfactor.cu", "tcfused.cu", "f64block.cu", "iterate.cu", "iterate_split.cu", "iterate_res.cu", "iterate_row.cu", "unroll.cu", "backward.cu", "lu.cu", "devtools.cu", "hostio.cu"]
HEADERS = ["common.cuh", "layout.cuh", "itergeom.cuh", "tcmma.cuh", os.path.join("..", "..", "include", "lqpb.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-warn-spills"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    extra = os.environ.get("LQPB_EXTRA_NVCC_FLAGS", "").split()
    if extra:           # developer builds (e.g. -DLQPB_PHASE_TIMERS) never reuse the cached library
        force = True
    dig = _digest()
    if not force and os.path.exists(SO) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return SO
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(PKG, "_build_" + src.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if out.strip() and (verbose or pr.returncode != 0 or "warning" in out.lower() or "spill" in out.lower()):
            print(f"--- nvcc {src}\n{out}", file=sys.stderr)
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    with open(STAMP, "w") as fh:
        fh.write(dig if not extra else "developer build: " + " ".join(extra))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
