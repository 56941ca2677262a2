/*
 * vgpu_b200.h - C ABI of libvgpu-control.so (B200-native build).
 *
 * PART 1 is the drop-in boundary: the symbols the dynamic linker / cudart / NVML clients bind
 * when the library is listed in /etc/ld.so.preload (reference Dockerfile:39-45,
 * pkg/deviceplugin/vgpu/vnum_plugin.go:715-746) or LD_PRELOAD (pkg/kubeletplugin/vgpu.go:176).
 * Signatures are the NVIDIA driver's own; each line cites the reference hook it replaces.
 * PART 2 are additional entry points (prefix vgpu_b200_) that expose the device kernels
 * directly - used by bench.py, the parity tests and operators; the reference has no such API.
 *
 * Plain C, pointers and sizes only.  Types below are ABI-identical stand-ins for the CUDA /
 * NVML typedefs so that this header can be included without <cuda.h>.
 */
#ifndef VGPU_B200_H
#define VGPU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef VGPU_B200_NO_DRIVER_TYPES
typedef int vb_CUresult;                 /* CUresult      */
typedef int vb_CUdevice;                 /* CUdevice      */
typedef unsigned long long vb_CUdeviceptr; /* CUdeviceptr */
typedef void *vb_handle;                 /* CUstream / CUfunction / CUarray / CUmemoryPool / nvmlDevice_t */
typedef int vb_nvmlReturn;               /* nvmlReturn_t  */
#endif

/* ======================================================================== PART 1: hook surface
 * (declared as comments-with-prototypes so including this header next to <cuda.h> never clashes
 * with the toolkit's own declarations; the exported symbol list is pinned by
 * tests/test_symbols.py against this table)
 *
 *  symbol                                   replaces (reference file:line)
 *  ---------------------------------------  ------------------------------------------------
 *  void *dlsym(void*, const char*)          library/src/loader.c:1517
 *  cuDriverGetVersion(int*)                 library/src/cuda_hook.c:1162
 *  cuInit(unsigned)                         library/src/cuda_hook.c:1176
 *  cuGetProcAddress(sym,pfn,ver,flags)      library/src/cuda_hook.c:1190
 *  cuGetProcAddress_v2(sym,pfn,ver,flags,st) library/src/cuda_hook.c:1262
 *  cuMemAllocManaged(dptr,bytes,flags)      library/src/cuda_hook.c:1316
 *  cuMemAlloc / cuMemAlloc_v2               library/src/cuda_hook.c:1344-1398
 *  cuMemAllocPitch / _v2                    library/src/cuda_hook.c:1400-1462
 *  cuMemAllocAsync / _ptsz                  library/src/cuda_hook.c:1464-1544
 *  cuArrayCreate / _v2                      library/src/cuda_hook.c:1572-1607
 *  cuArray3DCreate / _v2                    library/src/cuda_hook.c:1609-1644
 *  cuMipmappedArrayCreate                   library/src/cuda_hook.c:1646
 *  cuMemCreate                              library/src/cuda_hook.c:1672
 *  cuDeviceTotalMem / _v2                   library/src/cuda_hook.c:1699-1726
 *  cuMemGetInfo / _v2                       library/src/cuda_hook.c:1728-1808
 *  cuLaunchKernel / _ptsz                   library/src/cuda_hook.c:1810-1849
 *  cuLaunchKernelEx / _ptsz                 library/src/cuda_hook.c:1851-1881
 *  cuLaunch                                 library/src/cuda_hook.c:1883
 *  cuLaunchCooperativeKernel / _ptsz        library/src/cuda_hook.c:1901-1938
 *  cuLaunchGrid / cuLaunchGridAsync         library/src/cuda_hook.c:1940-1974
 *  cuFuncSetBlockShape                      library/src/cuda_hook.c:1976
 *  cuMemAllocFromPoolAsync / _ptsz          library/src/cuda_hook.c:2004-2050
 *  cuMemFree / _v2                          library/src/cuda_hook.c:2052-2079
 *  cuMemFreeAsync / _ptsz                   library/src/cuda_hook.c:2081-2109
 *  nvmlInit / nvmlInit_v2 / nvmlInitWithFlags  library/src/nvml_hook.c:32-45
 *  nvmlDeviceGetMemoryInfo                  library/src/nvml_hook.c:47
 *  nvmlDeviceGetMemoryInfo_v2               library/src/nvml_hook.c:75
 *  nvmlDeviceSetComputeMode                 library/src/nvml_hook.c:105
 *  nvmlDeviceGetPersistenceMode             library/src/nvml_hook.c:121
 *  nvmlDeviceGetUtilizationRates            library/src/nvml_originals.c:698 (forward)
 *  cuCtxSynchronize                         (none - B200 addition: asks a resident sampler
 *                                            kernel to retire before the tenant's device sync)
 *  cuStreamSynchronize / _ptsz              (none - B200 addition: blocking calls of a throttled tenant first
 *  cuEventSynchronize                         wait in user space while its work is parked behind the gate -
 *  cuMemcpyDtoH_v2 / _ptds                    where the reference's thread would be asleep in rate_limiter)
 *  cuMemcpyHtoD_v2 / _ptds                  (same: synchronous copies that may involve host memory wait
 *  cuMemcpy / _ptds                           for the legacy / per-thread stream inside the driver)
 *  cuStreamDestroy_v2                       (none - B200 addition: releases the stream's
 *                                            completion-marker slot)
 *  cuCtxDestroy / _v2                       (none - B200 addition: the token bucket, slab, streams and
 *  cuDevicePrimaryCtxReset / _v2              module live in the tenant's context; these detach the
 *  cuDevicePrimaryCtxRelease / _v2            runtime before the driver frees it, the next hooked call
 *                                            rebuilds it in the new context)
 *  cuGraphInstantiateWithFlags              library/src/cuda_originals.c:2953-2980 (forward) - B200
 *  cuGraphInstantiateWithParams / _ptsz       addition, opt-in VGPU_B200_GRAPH_LIMIT=1: learns the
 *  cuGraphExecDestroy                         token cost of a graph (sum of its kernel nodes' grids)
 *  cuGraphLaunch / _ptsz                    library/src/cuda_originals.c:3033 (forward) - opt-in:
 *                                            a replay pays that cost and is gated like a launch
 */

/* ======================================================================== PART 2: direct API */

/* Bring the device runtime up in the calling thread's current CUDA context (loads the embedded
 * sm_100a image, allocates the HBM token bucket and UVA slab).  0 on success, -1 if there is no
 * current context or the image cannot be loaded - there is no CPU fallback. */
int vgpu_b200_attach(void);
const char *vgpu_b200_version(void);

/* 128-bit vectorised clear of `bytes` at device address `dst`, enqueued on `stream`
 * (a CUstream; NULL = legacy default stream).  Returns a CUresult. */
int vgpu_b200_clear(unsigned long long dst, size_t bytes, void *stream);

/* Spill/stage copy of `bytes` from `src` to `dst` (device or managed addresses) with TMA bulk
 * copies.  Algorithmic traffic 2*bytes.  Returns a CUresult. */
int vgpu_b200_spill_copy(unsigned long long dst, unsigned long long src, size_t bytes, void *stream);

/* Run the memory-quota kernel on a caller-supplied request (host memory) and wait for the
 * answer.  Struct layouts: vgpu_manager_b200/csrc/kernel_abi.h.  0 / -1. */
struct vgpu_quota_req_s;
struct vgpu_quota_res_s;
int vgpu_b200_quota_eval(const void *req /* vgpu_quota_req_t */, void *res /* vgpu_quota_res_t */);

/* Device-resident UVA slab: insert returns 0 / -1 (table full); remove returns 0 (found, *bytes
 * set), 1 (not recorded) or -1 (error). */
int vgpu_b200_slab_insert(unsigned long long dptr, unsigned long long bytes);
int vgpu_b200_slab_remove(unsigned long long dptr, unsigned long long *bytes);

/* One operation on the slab placement table of the VGPU_B200_SLAB mode (vgpu_vslab_req_t ->
 * vgpu_vslab_res_t, kernel_abi.h): PUT = free-slot scan + insert, TAKE = lookup + remove, SCAN =
 * coldest slab of a size class in a given placement, its placement bits flipped in the same
 * launch (the spill / promote decision).  0 / -1. */
int vgpu_b200_vslab_op(const void *req /* vgpu_vslab_req_t */, void *res /* vgpu_vslab_res_t */);

typedef struct {
  long long granted;   /* cumulative grant (HBM)                                   */
  long long consumed;  /* cumulative consumption (host hook)                       */
  long long bucket;    /* granted - consumed after the last step == g_cur_cuda_cores */
  long long share;     /* == shares[] (reference cuda_hook.c:369)                  */
  int up_limit, sys_free, avg_sys_free, ctr_i, pre_sys_process_num, valid;
  int user_current, sys_current, sm_active_pct, queue_busy_pct;
  unsigned long long steps;
} vgpu_b200_limiter_state_t;

/* Re-initialise the HBM limiter state (sm_num <= 0 keeps the device's real SM geometry). */
int vgpu_b200_limiter_reset(int sm_num, int max_thread_per_sm, int hard_core, int soft_core,
                            int core_limit, int hard_limit);
/* One controller step on the device with an explicit utilisation reading
 * (== one iteration of reference cuda_hook.c:413-466). */
int vgpu_b200_limiter_step(int user_current, int sys_current, int valid, int sys_process_num,
                           vgpu_b200_limiter_state_t *out);
/* One default control step on a caller-supplied utilisation publication (vgpu_util_req_t,
 * kernel_abi.h): raw per-process samples + membership flags are folded on the device
 * (== reference get_used_gpu_utilization, cuda_hook.c:1044-1159) and the watcher body runs on
 * the result, in one launch of vgpu_refill_kernel.  0 / -1. */
int vgpu_b200_refill(const void *util_req /* vgpu_util_req_t */, vgpu_b200_limiter_state_t *out);
/* Host-side consumption of tokens (what a launch hook does), for tests. */
int vgpu_b200_limiter_consume(long long tokens);
int vgpu_b200_limiter_state(vgpu_b200_limiter_state_t *out);
/* One sampler launch (window/interval in microseconds; the controller runs when
 * `period_ticks` launches have accumulated).  user_override >= 0 replaces the measured
 * utilisation (test hook), -1 uses the measurement. */
int vgpu_b200_sampler_run(unsigned window_us, unsigned interval_us, unsigned period_ticks,
                          int user_override, vgpu_b200_limiter_state_t *out);

/* Node-level rebalance (no reference counterpart): assign tenant(s) of GPU
 * `host_index` a utilisation target `up_limit` (percent) under a ceiling `soft_core`; a ceiling
 * above the tenant's hard quota switches it to balance mode with that target.  Written to
 * /etc/vgpu-manager/config/rebalance.config, applied by each tenant at its next control step. */
int vgpu_b200_set_limits(int host_index, int up_limit, int soft_core);

/* Tuning: bytes per TMA bulk copy (multiple of 16), ring depth (2..16), CTAs per SM; the ring must
 * fit 200 KiB of shared memory.  0 / -1. */
int vgpu_b200_set_spill_geometry(unsigned chunk, unsigned stages, unsigned ctas_per_sm);
unsigned long long vgpu_b200_self_bytes(void);
unsigned long long vgpu_b200_metric(int host_index, int which);

#ifdef __cplusplus
}
#endif
#endif
