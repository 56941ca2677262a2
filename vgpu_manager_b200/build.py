"""Build the interception library in-tree (no torch, no setuptools):

    kernels.cu  --nvcc sm_100a-->  kernels.fatbin  --bin2c-->  kernels_image.gen.c
    *.c + kernels_image.gen.c  --gcc -shared-->  vgpu_manager_b200/libvgpu-control.so

The host side is plain C (like the reference's library/CMakeLists.txt: -O2 -g, libc only); the
device side is compiled for sm_100a only.  `python -m vgpu_manager_b200.build` or
`__graft_entry__.build()`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvgpu-control.so")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
BIN2C = shutil.which("bin2c") or "/usr/local/cuda/bin/bin2c"
HOST_SRCS = ["boot.c", "hooktab.c", "config.c", "device.c", "memgate.c", "slabmode.c", "limiter.c", "lifecycle.c", "metrics.c"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout


def build(force=False, verbose=False):
    fatbin = os.path.join(CSRC, "kernels.fatbin")
    image_c = os.path.join(CSRC, "kernels_image.gen.c")
    cu = os.path.join(CSRC, "kernels.cu")
    abi = os.path.join(CSRC, "kernel_abi.h")
    contract = os.path.join(HERE, "..", "include", "vgpu_contract.h")
    if force or _newer(fatbin, [cu, abi, contract]):
        _run([NVCC, *NVCC_FLAGS, "-fatbin", "-o", fatbin, cu])
    if force or _newer(image_c, [fatbin]):
        body = _run([BIN2C, "--const", "--name", "vgpu_kernels_image", fatbin])
        with open(image_c, "w") as f:
            f.write("/* generated from kernels.cu by vgpu_manager_b200/build.py - do not edit */\n")
            f.write(body)
            f.write("\nconst unsigned long long vgpu_kernels_image_size = sizeof(vgpu_kernels_image);\n")
    srcs = [os.path.join(CSRC, s) for s in HOST_SRCS] + [image_c]
    hdrs = [os.path.join(CSRC, h) for h in ("vgpu_internal.h", "kernel_abi.h", "cu_abi_subset.h")] + [
        contract, os.path.join(HERE, "..", "include", "vgpu_b200.h")]
    if force or _newer(OUT, srcs + hdrs):
        cmd = ["gcc", "-D_GNU_SOURCE", "-std=gnu11", "-O2", "-g", "-Wall", "-Wshadow", "-fPIC", "-shared",
               "-fvisibility=hidden", "-pthread", "-static-libgcc", "-o", OUT, *srcs, "-ldl"]
        out = _run(cmd)
        if verbose and out:
            print(out)
    build_tools(force)
    return OUT


WATCHER = os.path.join(HERE, "vgpu-smwatcher")


def build_tools(force=False):
    """Node-level helper binaries (plain C, libc only)."""
    src = os.path.join(CSRC, "smwatcher.c")
    contract = os.path.join(HERE, "..", "include", "vgpu_contract.h")
    if force or _newer(WATCHER, [src, contract]):
        _run(["gcc", "-D_GNU_SOURCE", "-std=gnu11", "-O2", "-g", "-Wall", "-Wshadow", "-o", WATCHER, src, "-ldl"])
    return WATCHER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
