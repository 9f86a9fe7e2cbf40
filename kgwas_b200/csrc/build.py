"""Build libkgwas_b200.so in-tree with nvcc for sm_100a (the only target)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(os.path.dirname(HERE), "libkgwas_b200.so")
SOURCES = ["kgb_api.cu", "kgb_csr.cu", "kgb_spmm.cu", "kgb_gemm_ffma.cu", "kgb_gemm_tc.cu", "kgb_elem.cu", "kgb_gat.cu", "kgb_sampler.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"), "-I", HERE]


def _obj(src):
    return os.path.join(HERE, "build", src.replace(".cu", ".o"))


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [os.path.join(HERE, src), os.path.join(ROOT, "include", "kgwas_b200.h")]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    procs = []
    for s in srcs:
        o = _obj(s)
        if force or _stale(s, o):
            cmd = ["nvcc", *FLAGS, "-c", os.path.join(HERE, s), "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0):
            print(f"--- {s}\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(OUT):
        cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *[_obj(s) for s in srcs]]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
