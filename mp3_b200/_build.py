"""Build recipes: nvcc (sm_100a) for libmp3b.so, gcc for the generator.  In-tree outputs so the
.so files travel to the GPU box with the repo snapshot."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libmp3b.so")

CU_SOURCES = ["k_index.cu", "k_huffman.cu", "k_requant.cu", "k_hybrid.cu", "k_synth.cu", "k_fused.cu", "k_resample.cu", "k_resample_tc.cu", "k_stretch.cu", "k_layer2.cu", "k_planar.cu", "k_segments.cu",
              "api.cu"]
CPP_SOURCES = ["tables_build.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(ROOT, "include", "mp3b.h"))
    return hs


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link mp3_b200/libmp3b.so."""
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    objs = []
    rebuilt = False
    for src in CU_SOURCES + CPP_SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        op = os.path.join(OBJ, src + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + hdrs):
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", op]
            subprocess.check_call(cmd)
            rebuilt = True
    if rebuilt or force or _stale(LIB, objs):
        subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
                              + ["-lcudart"])
    return LIB


def build_all(force=False):
    from . import synth
    synth.build(force)
    return build_lib(force)


def build_variant(name, defines=(), sources=None):
    """A tuning variant of the library: the translation units in `sources` (default: the two hot kernels) are
    recompiled with extra -D flags and linked with the main build's other objects into
    mp3_b200/variants/libmp3b_<name>.so.  Selected at run time with MP3B_LIB (tools/variant_bench.py)."""
    build_lib()
    sources = list(sources or ["k_fused.cu", "k_huffman.cu"])
    vdir = os.path.join(HERE, "variants")
    odir = os.path.join(vdir, "obj_" + name)
    os.makedirs(odir, exist_ok=True)
    objs = []
    for src in CU_SOURCES + CPP_SOURCES:
        if src in sources:
            op = os.path.join(odir, src + ".o")
            subprocess.check_call([_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", op])
        else:
            op = os.path.join(OBJ, src + ".o")
        objs.append(op)
    lib = os.path.join(vdir, "libmp3b_%s.so" % name)
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lcudart"])
    return lib
