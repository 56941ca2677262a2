"""CPU: the C-ABI boundary - exported symbol surface, libc-only linkage, ABI type layout."""
import os
import re
import subprocess

import helpers as H

ROOT = H.ROOT
# reference library/src/cuda_hook.c:243-281 (37 names) and library/src/nvml_hook.c:20-28 (7 names)
REFERENCE_CUDA_HOOKS = """cuDriverGetVersion cuInit cuGetProcAddress cuGetProcAddress_v2 cuMemAllocManaged cuMemAlloc_v2
cuMemAlloc cuMemAllocPitch_v2 cuMemAllocPitch cuArrayCreate_v2 cuArrayCreate cuArray3DCreate_v2 cuArray3DCreate
cuMipmappedArrayCreate cuDeviceTotalMem_v2 cuDeviceTotalMem cuMemGetInfo_v2 cuMemGetInfo cuLaunchKernel_ptsz
cuLaunchKernel cuLaunchKernelEx_ptsz cuLaunchKernelEx cuLaunch cuLaunchCooperativeKernel_ptsz cuLaunchCooperativeKernel
cuLaunchGrid cuLaunchGridAsync cuFuncSetBlockShape cuMemAllocAsync cuMemAllocAsync_ptsz cuMemCreate
cuMemAllocFromPoolAsync cuMemAllocFromPoolAsync_ptsz cuMemFree_v2 cuMemFree cuMemFreeAsync cuMemFreeAsync_ptsz""".split()
REFERENCE_NVML_HOOKS = """nvmlInit nvmlInit_v2 nvmlInitWithFlags nvmlDeviceGetMemoryInfo nvmlDeviceGetMemoryInfo_v2
nvmlDeviceSetComputeMode nvmlDeviceGetPersistenceMode""".split()


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {l.split()[-1] for l in out.splitlines() if l.strip()}


def test_exports_the_reference_hook_surface_and_nothing_accidental(built):
    sym = exported(H.NEW_SO)
    assert len(REFERENCE_CUDA_HOOKS) == 37 and len(REFERENCE_NVML_HOOKS) == 7
    missing = [s for s in REFERENCE_CUDA_HOOKS + REFERENCE_NVML_HOOKS + ["dlsym"] if s not in sym]
    assert not missing, missing
    extra = {s for s in sym if not s.startswith("vgpu_b200_")} - set(REFERENCE_CUDA_HOOKS) - set(REFERENCE_NVML_HOOKS)
    graph_opt_in = {"cuGraphInstantiateWithFlags", "cuGraphInstantiateWithParams", "cuGraphInstantiateWithParams_ptsz",
                    "cuGraphExecDestroy", "cuGraphLaunch", "cuGraphLaunch_ptsz"}
    ctx_teardown = {"cuCtxDestroy", "cuCtxDestroy_v2", "cuDevicePrimaryCtxReset", "cuDevicePrimaryCtxReset_v2",
                    "cuDevicePrimaryCtxRelease", "cuDevicePrimaryCtxRelease_v2"}
    blocking = {"cuStreamSynchronize", "cuStreamSynchronize_ptsz", "cuEventSynchronize", "cuMemcpyDtoH_v2", "cuMemcpyDtoH_v2_ptds",
                "cuMemcpyHtoD_v2", "cuMemcpyHtoD_v2_ptds", "cuMemcpy", "cuMemcpy_ptds"}
    assert extra == {"dlsym", "cuCtxSynchronize", "cuStreamDestroy_v2", "nvmlDeviceGetUtilizationRates"} | graph_opt_in | ctx_teardown | blocking, extra
    if H.have_reference():
        ref = exported(H.REF_SO)
        assert set(REFERENCE_CUDA_HOOKS + REFERENCE_NVML_HOOKS + ["dlsym"]) <= ref


def test_every_symbol_declared_in_the_public_header_is_exported(built):
    sym = exported(H.NEW_SO)
    with open(os.path.join(ROOT, "include", "vgpu_b200.h")) as f:
        text = f.read()
    part2 = text[text.index("PART 2: direct API"):]
    declared = set(re.findall(r"\b(vgpu_b200_\w+)\s*\(", part2))
    assert declared and declared <= sym, declared - sym
    part1 = text[text.index("PART 1: hook surface"):text.index("PART 2: direct API")]
    names = set()
    for line in part1.splitlines():
        m = re.match(r"\s*\*\s+((?:cu|nvml)[\w /]+?)(?:\(|\s{2,})", line)
        if not m:
            continue
        parts = [p.strip() for p in m.group(1).split("/")]
        base = parts[0]
        names.add(base)
        for p in parts[1:]:
            names.add(p if p.startswith(("cu", "nvml")) else base + p)
    assert len(names) >= 40
    assert names <= sym, names - sym


def test_links_only_libc_like_the_reference(built):
    out = subprocess.run(["readelf", "-d", H.NEW_SO], capture_output=True, text=True, check=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert all(n.startswith(("libc.so", "ld-linux", "libdl", "libpthread")) for n in needed), needed


def test_embedded_image_is_sm_100a_only(built):
    fatbin = os.path.join(ROOT, "vgpu_manager_b200", "csrc", "kernels.fatbin")
    out = subprocess.run(["cuobjdump", "-lelf", fatbin], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "vgpu_spill_copy_kernel", fatbin], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass, "the spill copy must be a TMA bulk copy (UBLKCP in SASS)"
    assert "SYNCS" in sass  # mbarrier traffic
    clear = subprocess.run(["cuobjdump", "-sass", "-fun", "vgpu_clear_kernel", fatbin], capture_output=True, text=True).stdout
    assert "STG.E" in clear and ".128" in clear, "the clear must use 128-bit stores"


def test_restated_driver_types_match_cuda_12_9_headers(built):
    src = os.path.join(ROOT, "tests", "harness", "abi_layout.c")
    r = subprocess.run(["gcc", "-std=gnu11", "-D_GNU_SOURCE", "-I/usr/local/cuda/include", "-fsyntax-only", src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure; the shipped library must not include or link it."""
    csrc = os.path.join(ROOT, "vgpu_manager_b200")
    for dirpath, _, files in os.walk(csrc):
        for fn in files:
            if fn.endswith((".c", ".h", ".cu", ".py")):
                with open(os.path.join(dirpath, fn), errors="ignore") as f:
                    t = f.read()
                assert "vgpu_oracle" not in t and "liboracle" not in t and "orc_" not in t, fn
