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


def build_omp(force=False):
    """C++/OpenMP port of the same sources (-DKNP_EMU -fopenmp -O3): the CPU baseline that bench.py
    times on the GPU box's host cores (`cpu_baseline`, `--impl reference`).  Lives under oracle/
    because it is measurement infrastructure, never loaded by the package."""
    gen = generate_models()
    outdir = os.path.join(ROOT, "oracle", "_port")
    os.makedirs(outdir, exist_ok=True)
    out = os.path.join(outdir, "libknpemi_omp.so")
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [gen, os.path.join(ROOT, "include", "knpemi.h")]
    if not force and not _stale(out, deps):
        return out
    cmd = ["g++", "-std=c++17", "-O3", "-fopenmp", "-DKNP_EMU", "-fPIC", "-pthread", "-shared", "-o", out]
    for f in SOURCES:
        cmd += ["-x", "c++", os.path.join(CSRC, f)]
    subprocess.run(cmd, check=True)
    return out


def build_variant(user_modules, emu=False, cache_dir=None):
    """A copy of the library that carries, besides the bundled membrane models, the given
    user ODE modules (anything that follows the reference's mm_*.py protocol: the Python source
    of `rhs_numba` is translated to a device function by knpemidg.odegen).  Returns
    (path of the .so, {module: model name}).  Cached by content."""
    import hashlib
    sys.path.insert(0, HERE)
    from knpemidg import odegen
    from knpemidg.models import BUNDLED
    mods = [(name, importlib.import_module("knpemidg.models." + name)) for name in BUNDLED]
    names = {}
    for m in user_modules:
        src = odegen.python_rhs_source(m)
        tables = repr((list(m.init_state_values()), list(m.init_parameter_values())))
        tag = hashlib.sha1((src + tables).encode()).hexdigest()[:10]
        base = "".join(ch if ch.isalnum() else "_" for ch in m.__name__.split(".")[-1])
        names[m] = f"user_{base}_{tag}"
        mods.append((names[m], m))
    text = odegen.models_header(mods)
    text = text.replace('#include "../knp_ode.h"', f'#include "{os.path.join(CSRC, "knp_ode.h")}"')
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "knpemi.h")]
    stamp = hashlib.sha1((text + "".join(open(d).read() for d in deps) + ("emu" if emu else "cuda")).encode()).hexdigest()[:16]
    cache_dir = cache_dir or os.environ.get("KNPEMIDG_VARIANT_DIR") or os.path.join(HERE, "knpemidg", "_variants")
    vdir = os.path.join(cache_dir, stamp)
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, "libknpemi_emu.so" if emu else "libknpemi.so")
    if os.path.exists(out):
        return out, names
    header = os.path.join(vdir, "models_gen.h")
    with open(header, "w") as f:
        f.write(text)
    define = f'-DKNP_MODELS_HEADER="{header}"'
    tmp = out + f".tmp{os.getpid()}"     # per process: several ranks may build the same variant at once
    if emu:
        cmd = ["g++", "-std=c++17", "-O2", "-DKNP_EMU", define, "-fPIC", "-pthread", "-shared", "-o", tmp]
        for f in SOURCES:
            cmd += ["-x", "c++", os.path.join(CSRC, f)]
    else:
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc, "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", define,
               "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared", "--expt-relaxed-constexpr",
               "-I", nccl_include(), "-o", tmp] + [os.path.join(CSRC, f) for f in SOURCES] + ["-ldl"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, out)
    return out, names


if __name__ == "__main__":
    if "--omp" in sys.argv:
        print(build_omp(force="--force" in sys.argv))
    elif "--emu" in sys.argv:
        print(build_emu(force="--force" in sys.argv))
    else:
        print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
