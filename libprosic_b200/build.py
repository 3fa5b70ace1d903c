"""Build libvlr_b200.so in-tree with nvcc for sm_100a (no JIT cache; the .so travels with the repo)."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvlr_b200.so")
OBJ = os.path.join(HERE, "_build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "--expt-relaxed-constexpr"] + \
    os.environ.get("VLR_NVCC_DEFS", "").split()   # experiments only (e.g. -DVLR_OCC5=5); part of the build stamp


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def source_stamp():
    """SHA-1 over csrc/*, include/vlr_gpu.h and the compiler flags: compiled into the library
    (vlr_build_stamp) and compared at load time."""
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + ["../../include/vlr_gpu.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = source_stamp()
    if (not force and os.path.exists(OUT) and os.path.exists(stamp_file)
            and open(stamp_file).read() == stamp):
        return OUT
    if not os.path.exists(NVCC):
        if os.path.exists(OUT):
            return OUT  # GPU box without nvcc: the prebuilt library; _ffi.load() checks its build stamp
        raise RuntimeError("nvcc not found and no prebuilt libvlr_b200.so")
    os.makedirs(OBJ, exist_ok=True)

    def cc(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              (['-DVLR_BUILD_STAMP="%s"' % stamp] if src == "ctx.cu" else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, _sources()))
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    open(stamp_file, "w").write(stamp)
    return OUT


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
