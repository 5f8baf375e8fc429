"""Build libknpemi.so (sm_100a) and, on request, the host-emulation library
used only by the CPU test-suite.

    python knp-emi-dg_b200/build.py            # product: knpemidg/libknpemi.so
    python knp-emi-dg_b200/build.py --emu      # tests/emu/libknpemi_emu.so (g++, no CUDA)
"""
import importlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["knp_api.cu", "knp_solve.cu"]
HEADERS = ["knp_common.h", "knp_comm.h", "knp_dg.h", "knp_ode.h", "knp_linalg.h", "knp_amg.h", "knp_ctx.h"]


def generate_models():
    sys.path.insert(0, HERE)
    from knpemidg import odegen
    from knpemidg.models import BUNDLED
    mods = [(name, importlib.import_module("knpemidg.models." + name)) for name in BUNDLED]
    text = odegen.models_header(mods)
    gen = os.path.join(CSRC, "generated")
    os.makedirs(gen, exist_ok=True)
    path = os.path.join(gen, "models_gen.h")
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return path


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def nccl_include():
    """nccl.h (types only: the library resolves NCCL's functions from the already loaded
    libnccl.so.2 at run time, see csrc/knp_comm.h)"""
    try:
        import nvidia.nccl as pkg
        for base in list(getattr(pkg, "__path__", [])):
            inc = os.path.join(base, "include")
            if os.path.exists(os.path.join(inc, "nccl.h")):
                return inc
    except Exception:
        pass
    for inc in ("/usr/include", "/usr/local/cuda/include"):
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return inc
    raise RuntimeError("nccl.h not found")


def build_cuda(force=False, verbose=False):
    gen = generate_models()
    out = os.path.join(HERE, "knpemidg", "libknpemi.so")
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [gen, os.path.join(ROOT, "include", "knpemi.h")]
    if not force and not _stale(out, deps):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared", "--expt-relaxed-constexpr", "-I", nccl_include(),
           "-o", out] + [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return out


def build_emu(force=False):
    gen = generate_models()
    outdir = os.path.join(ROOT, "tests", "emu")
    os.makedirs(outdir, exist_ok=True)
    out = os.path.join(outdir, "libknpemi_emu.so")
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [gen, os.path.join(ROOT, "include", "knpemi.h")]
    if not force and not _stale(out, deps):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-DKNP_EMU", "-fPIC", "-pthread", "-shared", "-o", out]
    for f in SOURCES:
        cmd += ["-x", "c++", os.path.join(CSRC, f)]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force="--force" in sys.argv))
    else:
        print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
