"""Build fabind_b200/libfabind_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfabind_b200.so")
SOURCES = ["gemm.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_tc2.cu", "gemm_tc3.cu", "gemm_tc4.cu", "gemm_tc5.cu", "graph.cu", "layers.cu", "plus.cu", "forward.cu", "l2ops.cu", "postopt.cu", "backward.cu", "xatt_tc.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _deps():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [
        os.path.join(os.path.dirname(HERE), "include", "fabind_b200.h")]


def source_hash():
    """63-bit digest of the header + csrc/ (None when the sources are not shipped next to the library)"""
    import hashlib
    h = hashlib.sha256()
    try:
        for d in _deps():
            h.update(os.path.basename(d).encode())
            with open(d, "rb") as f:
                h.update(f.read())
    except OSError:
        return None
    return int.from_bytes(h.digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    if any(os.path.getmtime(d) > t for d in _deps()):
        return True
    stamp = LIB + ".hash"
    try:
        return int(open(stamp).read().strip()) != source_hash()
    except (OSError, ValueError):
        return True


def build(force=False, verbose=False, diag=False):
    """diag=True adds -DFB_DIAG: the diagnostic environment knobs (FB_PDL, FB_SKIP_CATS, FB_TC4, ...) exist only in such builds"""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + FLAGS + [f"-DFB_SOURCE_HASH={source_hash()}LL"] + (["-DFB_DIAG"] if diag else []) + os.environ.get("FB_EXTRA_NVCC_FLAGS", "").split() + \
              ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                  "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(LIB + ".hash", "w") as f:
        f.write(str(source_hash()))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, diag="--diag" in sys.argv))
