#!/usr/bin/env python3
"""Build librmx.so (sm_100a) in-tree:  python radio_mapper_b200/csrc/build.py [--force]

Each .cu is compiled in parallel with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17
and linked into radio_mapper_b200/librmx.so.  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "librmx.so")
OBJ = os.path.join(HERE, "_obj")
SOURCES = ["rmx_lib.cu", "rmx_inst_contig.cu", "rmx_inst_col.cu", "rmx_inst_contig32.cu", "rmx_inst_col32.cu", "rmx_inst_contig8.cu"]
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "rmx.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-diag-suppress", "177"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    deps = [os.path.join(HERE, src)] + [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not _stale(obj, deps):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-Xptxas", "-v", "-c", os.path.join(HERE, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build_variant(name, defines, verbose=False):
    """Developer aid: build radio_mapper_b200/_variants/librmx_<name>.so with extra -D flags (A/B runs
    through RMX_LIB_PATH).  Objects go to a per-variant directory."""
    global OBJ, OUT, FLAGS
    saved = (OBJ, OUT, FLAGS)
    try:
        vdir = os.path.join(PKG, "_variants")
        os.makedirs(vdir, exist_ok=True)
        OBJ = os.path.join(HERE, "_obj_" + name)
        OUT = os.path.join(vdir, "librmx_%s.so" % name)
        FLAGS = FLAGS + ["-D" + d for d in defines]
        return build(force=False, verbose=verbose)
    finally:
        OBJ, OUT, FLAGS = saved


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.unlink(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(_compile, SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    if force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + \
              ["-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:          # python build.py --variant NAME -DX=1 -DY=2
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a[2:] for a in sys.argv[i + 2:] if a.startswith("-D")], verbose="-v" in sys.argv))
    else:
        path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(path)
