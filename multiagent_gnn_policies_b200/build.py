"""Build libfgnn.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m multiagent_gnn_policies_b200.build [--force] [-v]

The heavy fused kernels are instantiated once per (K, HP) pair in their own object file so the
objects compile in parallel and only stale ones are rebuilt.  The .so / .o files are git-ignored
but travel with the gpurun snapshot.  No GPU is needed to build (nvcc cross-compiles).
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfgnn.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

KS = (1, 2, 3, 4)
HPS = (16, 32, 64, 128)

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--fmad=true"] + os.environ.get("FGNN_NVCC_FLAGS", "").split()


def nvcc_path():
    cand = os.environ.get("NVCC")
    if cand:
        return cand
    return "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"


def _units():
    """(object path, source, extra flags, dependency list)"""
    common_deps = [os.path.join(CSRC, "fgnn_kernels.cuh")]
    units = [(os.path.join(OBJ, "fgnn.o"), os.path.join(CSRC, "fgnn.cu"), [],
              common_deps + [os.path.join(INCLUDE, "fgnn.h"), os.path.join(CSRC, "fgnn_final.cuh"),
                             os.path.join(CSRC, "fgnn_final_tc.cuh"), os.path.join(CSRC, "fgnn_pair.cuh")])]
    units.append((os.path.join(OBJ, "fgnn_train.o"), os.path.join(CSRC, "fgnn_train.cu"), [],
                  [os.path.join(INCLUDE, "fgnn.h")]))
    units.append((os.path.join(OBJ, "fgnn_dense.o"), os.path.join(CSRC, "fgnn_dense.cu"), [],
                  [os.path.join(INCLUDE, "fgnn.h")]))
    for k in KS:
        for hp in HPS:
            units.append((os.path.join(OBJ, f"fgnn_final_k{k}_hp{hp}.o"), os.path.join(CSRC, "fgnn_final.cu"),
                          [f"-DFGNN_K={k}", f"-DFGNN_HP={hp}"],
                          common_deps + [os.path.join(CSRC, "fgnn_final.cuh"), os.path.join(CSRC, "fgnn_final_tc.cuh")]))
            units.append((os.path.join(OBJ, f"fgnn_mini_k{k}_hp{hp}.o"), os.path.join(CSRC, "fgnn_mini.cu"),
                          [f"-DFGNN_K={k}", f"-DFGNN_HP={hp}"],
                          common_deps + [os.path.join(CSRC, "fgnn_final.cuh"), os.path.join(CSRC, "fgnn_final_tc.cuh")]))
    return units


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(unit, verbose):
    obj, src, flags, _ = unit
    cmd = [nvcc_path()] + COMMON + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    units = _units()
    todo = [u for u in units if force or _stale(u[0], [u[1]] + u[3])]
    if todo:
        print(f"[fgnn build] compiling {len(todo)} object(s) for sm_100a with {nvcc_path()}", flush=True)
        workers = min(len(todo), os.cpu_count() or 4)
        # heaviest first so the pool drains evenly
        todo.sort(key=lambda u: ("hp64" in u[0], "hp128" in u[0], "hp32" in u[0]), reverse=True)
        with concurrent.futures.ThreadPoolExecutor(workers) as pool:
            for obj, rc, out in pool.map(lambda u: _compile(u, verbose), todo):
                if verbose or rc:
                    print(out)
                if rc:
                    raise RuntimeError("nvcc failed for " + obj)
    objs = [u[0] for u in units]
    if todo or _stale(LIB, objs):
        cmd = [nvcc_path()] + ARCH + ["-shared", "-o", LIB] + objs
        print("[fgnn build] link", LIB, flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
