"""Builds librl_b200.so (hand-written sm_100a kernels + C-ABI) in-tree with nvcc.

`python -m rapid_locomotion_rl_b200.build` or `build()`; the .so is git-ignored but travels
with the working tree to the GPU box.  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "librl_b200.so")
STAMP = os.path.join(PKG, ".librl_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
] + os.environ.get("RL_NVCC_EXTRA", "").split()      # development switches (e.g. -DRL_CHAIN_TRACE_WRITE)
# env-path kernels mirror the eager reference op by op: no FMA contraction
NO_FMA = {"env_step.cu", "env_step_quad.cu", "env_step_rows.cu", "heights.cu", "env_reset.cu", "gac.cu", "gae.cu", "history.cu", "api.cu"}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for d, names in ((CSRC, sorted(os.listdir(CSRC))), (INCLUDE, sorted(os.listdir(INCLUDE)))):
        for n in names:
            if n.endswith((".cu", ".cuh", ".h")):  # noqa
                h.update(n.encode())
                with open(os.path.join(d, n), "rb") as f:
                    h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(" ".join(sorted(NO_FMA)).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _digest()


def build(force=False, verbose=False):
    if not force and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, src[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-fmad=false"] if src in NO_FMA else []) + \
              ["-I", INCLUDE, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose:
            sys.stderr.write(out)
        objs.append(obj)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
