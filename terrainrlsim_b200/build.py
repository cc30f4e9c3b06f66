"""In-tree build of libtrl_b200.so (nvcc, sm_100a). Used by __graft_entry__.build() and on first import."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtrl_b200.so")
SOURCES = ["trl_model.cpp", "trl_api.cpp", "trl_kernels.cu"]
STRICT_SOURCES = ["trl_terrain.cu"]  # compiled with -fmad=false: bit-exact float / double expression sequences
HEADERS = ["trl_internal.h", "trl_device.cuh", "trl_pd_aba.cuh", "trl_pd_block.cuh", os.path.join("..", "..", "include", "trl_b200.h")]


def nvcc_path():
    p = shutil.which("nvcc")
    if p:
        return p
    p = "/usr/local/cuda/bin/nvcc"
    return p if os.path.exists(p) else None


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in SOURCES + STRICT_SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into terrainrlsim_b200/lib/libtrl_b200.so."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libtrl_b200.so (there is no CPU fallback)")
    os.makedirs(LIB_DIR, exist_ok=True)
    common = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-diag-suppress", "177,128"]
    if verbose:
        common[1:1] = ["-Xptxas", "-v"]
    objs = []
    for src in STRICT_SOURCES:
        obj = os.path.join(LIB_DIR, os.path.splitext(src)[0] + ".o")
        subprocess.check_call(common + ["-fmad=false", "-c", src, "-o", obj], cwd=CSRC)
        objs.append(obj)
    # build to a temporary name and rename: ranks that import concurrently never see a half-written library
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    subprocess.check_call(common + ["-shared", "-o", tmp] + SOURCES + objs, cwd=CSRC)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


def build_diagnostic():
    """libtrl_b200_clk.so: the same library with -DTRL_PD_CLK (clock64 stamps of one k_pd_aba tile, read by
    tools/pd_phases.py through TRL_LIB=terrainrlsim_b200/lib/libtrl_b200_clk.so). Not loaded by default."""
    build()
    nvcc = nvcc_path()
    out = os.path.join(LIB_DIR, "libtrl_b200_clk.so")
    objs = [os.path.join(LIB_DIR, os.path.splitext(src)[0] + ".o") for src in STRICT_SOURCES]
    subprocess.check_call([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
                           "-diag-suppress", "177,128", "-DTRL_PD_CLK", "-shared", "-o", out] + SOURCES + objs, cwd=CSRC)
    return out
