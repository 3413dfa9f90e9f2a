"""Build recipe for the GPU-side reference checker (TEST INFRASTRUCTURE, never shipped, never on the product path).

Compiles the reference's own CUDA extension *from the sources where they lie* under
/root/reference/gsplat/cuda (nothing is copied into this repo) into ``oracle/_ref/gsplat_ref_cuda.so``
with the flags the reference's loader uses (gsplat/cuda/_backend.py:176-185: ``-O3 -use_fast_math``,
arch autodetected -> sm_100 on a B200).  The resulting pybind module exposes the reference `_C` ops
(gsplat/cuda/ext.cpp:6-104); `tests/` load it on the GPU box to compare our kernels against the
reference's kernels on identical inputs, and `bench.py` times it in its `ref_cuda` leg as "the kernel to beat".

oracle/_ref/ is git-ignored but travels to the GPU box with the gpurun snapshot.
Only runs where /root/reference exists (the build container); on the GPU box the prebuilt .so is used.
"""
import glob
import os
import sys

REF = "/root/reference/gsplat/cuda"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAME = "gsplat_ref_cuda"


def build(verbose: bool = False) -> str:
    so = os.path.join(OUT, NAME + ".so")
    if os.path.exists(so):
        return so
    if not os.path.isdir(REF):
        raise FileNotFoundError(f"{REF} not present (expected on the GPU box); prebuilt {so} missing")
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    sources = (
        sorted(glob.glob(os.path.join(REF, "csrc/*.cu")))
        + sorted(glob.glob(os.path.join(REF, "csrc/*.cpp")))
        + [os.path.join(REF, "ext.cpp")]
    )
    load(
        name=NAME,
        sources=sources,
        extra_cflags=["-O3", "-Wno-attributes"],
        extra_cuda_cflags=["-O3", "-use_fast_math"],
        extra_include_paths=[os.path.join(REF, "include/"), os.path.join(REF, "csrc", "third_party", "glm")],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,
    )
    return so


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
