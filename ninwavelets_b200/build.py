"""In-tree build of libnwcwt.so for sm_100a (nvcc cross-compiles without a GPU).

The transform kernels are instantiated per (kernel group, precision) in their own .cu files
and compiled in parallel; nwcwt.cu holds the C ABI."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
_SUF = os.environ.get("NW_LIB_SUFFIX", "")   # experiment builds (e.g. knock-out timing variants) live beside the product
OBJ = os.path.join(HERE, "build" + _SUF)
LIB = os.path.join(HERE, "libnwcwt%s.so" % _SUF)
SOURCES = ["nwcwt.cu", "k_short_f32.cu", "k_short_f64.cu", "k_passA_f32.cu", "k_passA_f64.cu",
           "k_passB_f32.cu", "k_passB_f64.cu", "k_long2_f32_c0.cu", "k_long2_f32_c1.cu", "k_long2_f32_c2.cu",
           "k_long2_f32_c3.cu", "k_long2_f64.cu", "k_short2_f32.cu", "k_short2_f64.cu", "k_short3_f32.cu", "k_resample_f32.cu",
           "k_resample_f64.cu", "k_resample_vec_f32_p4_m2.cu", "k_resample_vec_f32_p4_m1.cu", "k_resample_vec_f32_p2_m2.cu",
           "k_resample_vec_f32_p2_m1.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"] + os.environ.get("NW_EXTRA_NVCC_FLAGS", "").split()


def _deps(src=None):
    """Files an object depends on: every header of csrc, the public header and its own .cu (all .cu files for the library)."""
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu") or src is None or f == src]
    d.append(os.path.join(os.path.dirname(HERE), "include", "nwcwt.h"))
    return d


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    newest = max(os.path.getmtime(d) for d in _deps(src))
    if os.path.isfile(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stdout))
    if "warning" in r.stdout:   # e.g. #20013-D (host constexpr called from device code) silently miscompiles
        sys.stderr.write("nvcc warnings in %s:\n%s\n" % (src, r.stdout))
    return obj, r.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.isfile(LIB) and not os.path.isdir(OBJ):   # a shipped library without its objects: trust its date
        if os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in _deps()):
            return LIB
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 2)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    if os.path.isfile(LIB) and all(os.path.getmtime(o) <= os.path.getmtime(LIB) for o, _ in results):
        return LIB                                                      # every object is older than the library
    cmd = [os.environ.get("NVCC", "nvcc"), "-shared", "-o", LIB] + [o for o, _ in results]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
