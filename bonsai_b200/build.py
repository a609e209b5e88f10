"""Builds libbonsai_b200.so (hand-written sm_100a kernels + the C ABI) in-tree with nvcc.

    python -m bonsai_b200.build [--force]

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbonsai_b200.so")
SOURCES = ["bns_kernels.cu", "bns_api.cu", "bns_pack.cpp"]
DEPS = SOURCES + ["bns_device.cuh", "bns_classify_u.cuh", "bns_kernels.h", "bns_host_util.h", "bns_pack.h", os.path.join("..", "..", "include", "bonsai_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-pthread", "-shared", "-ldl"]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False, out=None, defines=()):
    if out is None and not force and not needs_build():
        return LIB
    out = out or LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libbonsai_b200.so")
    if out == LIB:
        with open(os.path.join(HERE, "build.log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    return out


CLI = os.path.join(HERE, "bin", "bonsai")
CLI_SRC = os.path.join(CSRC, "cli", "bonsai_main.cpp")
CLI_DEPS = [CLI_SRC, os.path.join(HERE, "..", "include", "bonsai_b200", "bonsai.hpp"), os.path.join(HERE, "..", "include", "bonsai_b200.h")]


def build_cli(force=False):
    """g++ the `bonsai` CLI (host C++ over the C ABI), linked against the in-tree library with an $ORIGIN rpath."""
    build()
    if not force and os.path.exists(CLI) and all(os.path.getmtime(d) <= os.path.getmtime(CLI) for d in CLI_DEPS + [LIB]):
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-o", CLI, CLI_SRC, "-L" + HERE, "-lbonsai_b200", "-lz",
           "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath," + HERE]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("g++ failed building the bonsai CLI")
    return CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_cli(force="--force" in sys.argv))
