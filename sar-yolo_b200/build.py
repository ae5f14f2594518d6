"""Build libsarpost.so in-tree with nvcc for sm_100a (one translation unit, static cudart)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsarpost.so")
SOURCES = ["sarpost.cu"]
DEPS = ["sarpost.cu", "common.cuh", "k1_candidates.cuh", "k2_select_sort.cuh", "k4_nms.cuh", "k6_match.cuh", "k7_state_head.cuh", "host_ctx.inl",
        os.path.join("..", "..", "include", "sarpost.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fopenmp", "-shared", "-cudart", "static", "-lgomp"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """`out` / `defines`: side builds for experiments (e.g. the phase-profiling variant: out=libsarpost_prof.so,
    defines=("SARPOST_PHASE_PROF",), loaded through SARPOST_LIB_PATH); the product library is always LIB."""
    if out is None and not force and not is_stale():
        return LIB
    extra = os.environ.get("SARPOST_EXTRA_NVCC_FLAGS", "").split() + [f"-D{d}" for d in defines]
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return out or LIB


if __name__ == "__main__":
    import sys

    if "--prof" in sys.argv:
        print(build(out=os.path.join(HERE, "libsarpost_prof.so"), defines=("SARPOST_PHASE_PROF",)))
    else:
        print(build(force=True, verbose="-v" in sys.argv))
