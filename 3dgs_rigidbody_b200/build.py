"""In-tree build of librigidsplat.so (hand-written sm_100a CUDA behind the C ABI of include/rigidsplat.h).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "librigidsplat.so")
SOURCES = ["abi.cu", "project.cu", "project_bwd.cu", "isect.cu", "depth_order.cu", "sort.cu", "raster_fwd.cu", "raster_bwd.cu", "frame.cu", "sh.cu", "exchange.cu", "cgc.cu", "seghead.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-use_fast_math", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _stamp() -> str:
    h = hashlib.sha256()
    deps = sorted(os.listdir(CSRC)) + ["../../include/rigidsplat.h"]
    for f in deps:
        path = os.path.join(CSRC, f)
        if os.path.isfile(path):
            h.update(f.encode())
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library. Returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []
    procs = []
    for s in srcs:
        obj = os.path.join(LIBDIR, s.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            print(f"--- nvcc failed on {s} ---\n{out}", file=sys.stderr)
        elif (verbose or ptxas_info) and out.strip():
            print(f"--- {s} ---\n{out}")
    if failed:
        raise RuntimeError("nvcc failed; see messages above")
    link = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    if verbose:
        print(" ".join(link), flush=True)
    subprocess.run(link, check=True)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv))
