"""Build geometry_rl_b200/libgrl_b200.so (sm_100a only) with nvcc.  No torch headers are involved:
the library exposes the plain C ABI of include/grl_b200.h and is bound with ctypes."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(PKG, "libgrl_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(PKG), "include", "grl_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(name, verbose):
    src = os.path.join(HERE, name)
    obj = os.path.join(OBJ, name[:-3] + ".o")
    if not _stale(src, obj):
        return name, 0, ""
    p = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
    return name, p.returncode, p.stdout + p.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    names = sources()
    logs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(names))) as ex:
        for name, rc, log in ex.map(lambda n: _compile(n, verbose), names):
            logs.append((name, log))
            if rc != 0:
                sys.stderr.write(log)
                raise RuntimeError(f"nvcc failed on {name}")
    objs = [os.path.join(OBJ, n[:-3] + ".o") for n in names]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        p = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                            "-cudart", "static"], capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("link failed")
    if verbose:
        for name, log in logs:
            if log:
                print(f"==== {name}\n{log}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
