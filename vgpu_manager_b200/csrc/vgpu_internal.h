/*
 * vgpu_internal.h - private declarations of the B200 interception library.
 *
 * The host side is plain C over the CUDA driver / NVML C ABIs.  We deliberately do not include
 * <cuda.h>/<nvml.h>: their versioning macros (#define cuMemAlloc cuMemAlloc_v2 ...) collide
 * with the very symbols this library must export.  Only the handful of types the hooked entry
 * points touch are restated here (vcu_ / vnv_ prefix); tests/test_abi_layout.c static-asserts
 * them against the real CUDA 12.9 headers.
 */
#ifndef VGPU_INTERNAL_H
#define VGPU_INTERNAL_H

#include <inttypes.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/vgpu_contract.h"
#include "kernel_abi.h"

#include "cu_abi_subset.h"

/* ------------------------------------------------------------------ logging
 * Same line format as the reference (hook.h:316-327); FATAL exits the process. */
enum { VL_FATAL = 0, VL_ERROR, VL_WARNING, VL_INFO, VL_VERBOSE, VL_DETAIL };
int vgpu_log_level(void);
void vgpu_log_emit(int level, const char *file, int line, const char *fmt, ...)
    __attribute__((format(printf, 4, 5)));
#define VLOG(level, ...)                                                        \
  do {                                                                          \
    if ((level) <= vgpu_log_level()) vgpu_log_emit((level), __FILE__, __LINE__, __VA_ARGS__); \
    if ((level) == VL_FATAL) exit(1);                                           \
  } while (0)

#define VGPU_EXPORT __attribute__((visibility("default")))
#define likely(x) __builtin_expect(!!(x), 1)
#define unlikely(x) __builtin_expect(!!(x), 0)

/* ------------------------------------------------------------------ real entry points
 * X(name, return type, (args)) - resolved from libcuda.so.<ver> / libnvidia-ml.so.<ver>. */
#define VGPU_REAL_CUDA(X)                                                                       \
  X(cuInit, CUresult, (unsigned int))                                                           \
  X(cuDriverGetVersion, CUresult, (int *))                                                      \
  X(cuGetProcAddress, CUresult, (const char *, void **, int, cuuint64_t))                       \
  X(cuGetProcAddress_v2, CUresult, (const char *, void **, int, cuuint64_t, void *))            \
  X(cuGetErrorString, CUresult, (CUresult, const char **))                                      \
  X(cuDeviceGetCount, CUresult, (int *))                                                        \
  X(cuDeviceGet, CUresult, (CUdevice *, int))                                                   \
  X(cuDeviceGetAttribute, CUresult, (int *, int, CUdevice))                                     \
  X(cuDeviceGetUuid, CUresult, (CUuuid *, CUdevice))                                            \
  X(cuDeviceGetUuid_v2, CUresult, (CUuuid *, CUdevice))                                         \
  X(cuDeviceTotalMem, CUresult, (size_t *, CUdevice))                                           \
  X(cuDeviceTotalMem_v2, CUresult, (size_t *, CUdevice))                                        \
  X(cuCtxGetDevice, CUresult, (CUdevice *))                                                     \
  X(cuCtxGetCurrent, CUresult, (CUcontext *))                                                   \
  X(cuCtxSetCurrent, CUresult, (CUcontext))                                                     \
  X(cuCtxPushCurrent_v2, CUresult, (CUcontext))                                                 \
  X(cuCtxPopCurrent_v2, CUresult, (CUcontext *))                                                \
  X(cuCtxSynchronize, CUresult, (void))                                                         \
  X(cuDevicePrimaryCtxRetain, CUresult, (CUcontext *, CUdevice))                                \
  X(cuDevicePrimaryCtxGetState, CUresult, (CUdevice, unsigned int *, int *))                    \
  X(cuDevicePrimaryCtxRelease, CUresult, (CUdevice))                                            \
  X(cuDevicePrimaryCtxRelease_v2, CUresult, (CUdevice))                                         \
  X(cuDevicePrimaryCtxReset, CUresult, (CUdevice))                                              \
  X(cuDevicePrimaryCtxReset_v2, CUresult, (CUdevice))                                           \
  X(cuCtxDestroy, CUresult, (CUcontext))                                                        \
  X(cuCtxDestroy_v2, CUresult, (CUcontext))                                                     \
  X(cuMemAlloc, CUresult, (CUdeviceptr *, size_t))                                              \
  X(cuMemAlloc_v2, CUresult, (CUdeviceptr *, size_t))                                           \
  X(cuMemAllocManaged, CUresult, (CUdeviceptr *, size_t, unsigned int))                         \
  X(cuMemAllocPitch, CUresult, (CUdeviceptr *, size_t *, size_t, size_t, unsigned int))         \
  X(cuMemAllocPitch_v2, CUresult, (CUdeviceptr *, size_t *, size_t, size_t, unsigned int))      \
  X(cuMemAllocAsync, CUresult, (CUdeviceptr *, size_t, CUstream))                               \
  X(cuMemAllocAsync_ptsz, CUresult, (CUdeviceptr *, size_t, CUstream))                          \
  X(cuMemAllocFromPoolAsync, CUresult, (CUdeviceptr *, size_t, CUmemoryPool, CUstream))         \
  X(cuMemAllocFromPoolAsync_ptsz, CUresult, (CUdeviceptr *, size_t, CUmemoryPool, CUstream))    \
  X(cuMemCreate, CUresult,                                                                      \
    (CUmemGenericAllocationHandle *, size_t, const vcu_mem_alloc_prop_t *, unsigned long long)) \
  X(cuMemRelease, CUresult, (CUmemGenericAllocationHandle))                                     \
  X(cuMemAddressReserve, CUresult, (CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long)) \
  X(cuMemAddressFree, CUresult, (CUdeviceptr, size_t))                                          \
  X(cuMemMap, CUresult, (CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long)) \
  X(cuMemUnmap, CUresult, (CUdeviceptr, size_t))                                                \
  X(cuMemSetAccess, CUresult, (CUdeviceptr, size_t, const vcu_mem_access_desc_t *, size_t))     \
  X(cuMemGetAllocationGranularity, CUresult, (size_t *, const vcu_mem_alloc_prop_t *, int))     \
  X(cuEventCreate, CUresult, (CUevent *, unsigned int))                                         \
  X(cuEventRecord, CUresult, (CUevent, CUstream))                                               \
  X(cuEventSynchronize, CUresult, (CUevent))                                                    \
  X(cuEventElapsedTime, CUresult, (float *, CUevent, CUevent))                                  \
  X(cuEventDestroy_v2, CUresult, (CUevent))                                                     \
  X(cuArrayCreate, CUresult, (CUarray *, const vcu_array_desc_t *))                             \
  X(cuArrayCreate_v2, CUresult, (CUarray *, const vcu_array_desc_t *))                          \
  X(cuArray3DCreate, CUresult, (CUarray *, const vcu_array3d_desc_t *))                         \
  X(cuArray3DCreate_v2, CUresult, (CUarray *, const vcu_array3d_desc_t *))                      \
  X(cuMipmappedArrayCreate, CUresult,                                                           \
    (CUmipmappedArray *, const vcu_array3d_desc_t *, unsigned int))                             \
  X(cuMemFree, CUresult, (CUdeviceptr))                                                         \
  X(cuMemFree_v2, CUresult, (CUdeviceptr))                                                      \
  X(cuMemFreeAsync, CUresult, (CUdeviceptr, CUstream))                                          \
  X(cuMemFreeAsync_ptsz, CUresult, (CUdeviceptr, CUstream))                                     \
  X(cuMemGetInfo, CUresult, (size_t *, size_t *))                                               \
  X(cuMemGetInfo_v2, CUresult, (size_t *, size_t *))                                            \
  X(cuMemHostAlloc, CUresult, (void **, size_t, unsigned int))                                  \
  X(cuMemFreeHost, CUresult, (void *))                                                          \
  X(cuMemGetAddressRange_v2, CUresult, (CUdeviceptr *, size_t *, CUdeviceptr))                  \
  X(cuMemHostGetDevicePointer_v2, CUresult, (CUdeviceptr *, void *, unsigned int))              \
  X(cuMemsetD8_v2, CUresult, (CUdeviceptr, unsigned char, size_t))                              \
  X(cuMemcpyDtoH_v2, CUresult, (void *, CUdeviceptr, size_t))                                   \
  X(cuMemcpyHtoD_v2, CUresult, (CUdeviceptr, const void *, size_t))                             \
  X(cuMemcpyDtoHAsync_v2, CUresult, (void *, CUdeviceptr, size_t, CUstream))                    \
  X(cuModuleLoadData, CUresult, (CUmodule *, const void *))                                     \
  X(cuModuleGetFunction, CUresult, (CUfunction *, CUmodule, const char *))                      \
  X(cuModuleUnload, CUresult, (CUmodule))                                                       \
  X(cuFuncSetAttribute, CUresult, (CUfunction, int, int))                                       \
  X(cuStreamCreate, CUresult, (CUstream *, unsigned int))                                       \
  X(cuStreamCreateWithPriority, CUresult, (CUstream *, unsigned int, int))                      \
  X(cuCtxGetStreamPriorityRange, CUresult, (int *, int *))                                      \
  X(cuStreamSynchronize, CUresult, (CUstream))                                                  \
  X(cuStreamSynchronize_ptsz, CUresult, (CUstream))                                             \
  X(cuMemcpyDtoH_v2_ptds, CUresult, (void *, CUdeviceptr, size_t))                              \
  X(cuMemcpyHtoD_v2_ptds, CUresult, (CUdeviceptr, const void *, size_t))                        \
  X(cuMemcpy, CUresult, (CUdeviceptr, CUdeviceptr, size_t))                                     \
  X(cuMemcpy_ptds, CUresult, (CUdeviceptr, CUdeviceptr, size_t))                                \
  X(cuStreamDestroy_v2, CUresult, (CUstream))                                                   \
  X(cuStreamQuery, CUresult, (CUstream))                                                        \
  X(cuStreamIsCapturing, CUresult, (CUstream, int *))                                           \
  X(cuThreadExchangeStreamCaptureMode, CUresult, (int *))                                       \
  X(cuStreamWaitValue64_v2, CUresult, (CUstream, CUdeviceptr, cuuint64_t, unsigned int))        \
  X(cuStreamWaitValue64_v2_ptsz, CUresult, (CUstream, CUdeviceptr, cuuint64_t, unsigned int))   \
  X(cuStreamWriteValue64_v2, CUresult, (CUstream, CUdeviceptr, cuuint64_t, unsigned int))       \
  X(cuStreamWriteValue64_v2_ptsz, CUresult, (CUstream, CUdeviceptr, cuuint64_t, unsigned int))  \
  X(cuLaunchKernel, CUresult,                                                                   \
    (CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, \
     void **, void **))                                                                         \
  X(cuLaunchKernel_ptsz, CUresult,                                                              \
    (CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, \
     void **, void **))                                                                         \
  X(cuLaunchKernelEx, CUresult, (const vcu_launch_config_t *, CUfunction, void **, void **))    \
  X(cuLaunchKernelEx_ptsz, CUresult, (const vcu_launch_config_t *, CUfunction, void **, void **)) \
  X(cuLaunchCooperativeKernel, CUresult,                                                        \
    (CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, \
     void **))                                                                                  \
  X(cuLaunchCooperativeKernel_ptsz, CUresult,                                                   \
    (CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, \
     void **))                                                                                  \
  X(cuLaunch, CUresult, (CUfunction))                                                           \
  X(cuLaunchGrid, CUresult, (CUfunction, int, int))                                             \
  X(cuLaunchGridAsync, CUresult, (CUfunction, int, int, CUstream))                              \
  X(cuFuncSetBlockShape, CUresult, (CUfunction, int, int, int))                                 \
  X(cuGraphInstantiateWithFlags, CUresult, (CUgraphExec *, CUgraph, unsigned long long))         \
  X(cuGraphInstantiateWithParams, CUresult, (CUgraphExec *, CUgraph, void *))                    \
  X(cuGraphInstantiateWithParams_ptsz, CUresult, (CUgraphExec *, CUgraph, void *))               \
  X(cuGraphLaunch, CUresult, (CUgraphExec, CUstream))                                           \
  X(cuGraphLaunch_ptsz, CUresult, (CUgraphExec, CUstream))                                      \
  X(cuGraphExecDestroy, CUresult, (CUgraphExec))                                                \
  X(cuGraphGetNodes, CUresult, (CUgraph, CUgraphNode *, size_t *))                              \
  X(cuGraphNodeGetType, CUresult, (CUgraphNode, int *))                                         \
  X(cuGraphKernelNodeGetParams, CUresult, (CUgraphNode, void *))                                \
  X(cuGraphKernelNodeGetParams_v2, CUresult, (CUgraphNode, void *))                             \
  X(cuGraphChildGraphNodeGetGraph, CUresult, (CUgraphNode, CUgraph *))

#define VGPU_REAL_NVML(X)                                                                       \
  X(nvmlInit, nvmlReturn_t, (void))                                                             \
  X(nvmlInit_v2, nvmlReturn_t, (void))                                                          \
  X(nvmlInitWithFlags, nvmlReturn_t, (unsigned int))                                            \
  X(nvmlErrorString, const char *, (nvmlReturn_t))                                              \
  X(nvmlDeviceGetCount, nvmlReturn_t, (unsigned int *))                                         \
  X(nvmlDeviceGetCount_v2, nvmlReturn_t, (unsigned int *))                                      \
  X(nvmlDeviceGetHandleByIndex, nvmlReturn_t, (unsigned int, nvmlDevice_t *))                   \
  X(nvmlDeviceGetHandleByIndex_v2, nvmlReturn_t, (unsigned int, nvmlDevice_t *))                \
  X(nvmlDeviceGetIndex, nvmlReturn_t, (nvmlDevice_t, unsigned int *))                           \
  X(nvmlDeviceGetUUID, nvmlReturn_t, (nvmlDevice_t, char *, unsigned int))                      \
  X(nvmlDeviceGetComputeRunningProcesses, nvmlReturn_t, (nvmlDevice_t, unsigned int *, vgpu_proc_t *)) \
  X(nvmlDeviceGetGraphicsRunningProcesses, nvmlReturn_t, (nvmlDevice_t, unsigned int *, vgpu_proc_t *)) \
  X(nvmlDeviceGetComputeRunningProcesses_v3, nvmlReturn_t, (nvmlDevice_t, unsigned int *, vgpu_proc_v2_t *)) \
  X(nvmlDeviceGetGraphicsRunningProcesses_v3, nvmlReturn_t, (nvmlDevice_t, unsigned int *, vgpu_proc_v2_t *)) \
  X(nvmlDeviceGetProcessUtilization, nvmlReturn_t,                                              \
    (nvmlDevice_t, vgpu_util_sample_t *, unsigned int *, unsigned long long))                   \
  X(nvmlDeviceGetUtilizationRates, nvmlReturn_t, (nvmlDevice_t, vnv_utilization_t *))           \
  X(nvmlDeviceGetMemoryInfo, nvmlReturn_t, (nvmlDevice_t, vnv_memory_t *))                      \
  X(nvmlDeviceGetMemoryInfo_v2, nvmlReturn_t, (nvmlDevice_t, vnv_memory_v2_t *))                \
  X(nvmlDeviceSetComputeMode, nvmlReturn_t, (nvmlDevice_t, int))                                \
  X(nvmlDeviceGetPersistenceMode, nvmlReturn_t, (nvmlDevice_t, int *))

typedef struct {
#define X(name, ret, args) ret(*name) args;
  VGPU_REAL_CUDA(X)
  VGPU_REAL_NVML(X)
#undef X
} vgpu_real_t;

extern vgpu_real_t R; /* real entry points; NULL when the driver lacks the symbol */
typedef void *(*vgpu_dlsym_fn)(void *, const char *);
extern vgpu_dlsym_fn vgpu_real_dlsym;

/* A library call that conflicts with a tenant's ongoing stream capture invalidates the tenant's
 * graph (CUDA_ERROR_STREAM_CAPTURE_* = 900..908): never let that pass silently. */
#define VGPU_CAPCHK(call) vgpu_capchk((call), #call, __FILE__, __LINE__)
static inline CUresult vgpu_capchk(CUresult r, const char *what, const char *file, int line) {
  if (__builtin_expect((int)r >= 900 && (int)r <= 908, 0))
    vgpu_log_emit(VL_ERROR, file, line, "stream-capture conflict: %s returned %d", what, (int)r);
  return r;
}

/* Library-internal driver calls that CUDA classifies as "potentially unsafe" (cuStreamQuery,
 * cuStreamSynchronize ...) would invalidate a graph capture that ANOTHER tenant thread has open
 * in global mode (seen with torch.cuda.graph: CUDA_ERROR_STREAM_CAPTURE_UNSUPPORTED from the tick
 * thread's cuStreamQuery).  They run with this thread's capture-interaction mode set to relaxed. */
static inline int vgpu_capture_relax(void) {
  int mode = VCU_STREAM_CAPTURE_MODE_RELAXED;
  if (R.cuThreadExchangeStreamCaptureMode && R.cuThreadExchangeStreamCaptureMode(&mode) == CUDA_SUCCESS) return mode;
  return -1;
}
static inline void vgpu_capture_restore(int prev) {
  if (prev >= 0 && R.cuThreadExchangeStreamCaptureMode) R.cuThreadExchangeStreamCaptureMode(&prev);
}

/* ------------------------------------------------------------------ global state */
extern vgpu_cfg_t *G_cfg;       /* mmap'ed (RO) or env-built vgpu.config */
extern vgpu_smutil_t *G_smutil; /* optional external watcher file        */
extern vgpu_vmem_t *G_vmem;     /* shared UVA ledger file (RW)           */

/* Optional development/test sandbox: VGPU_B200_SANDBOX=/dir prefixes every contract path
 * (never set in production; the control plane knows nothing about it). */
const char *vgpu_path(const char *abs, char *buf, size_t cap);
const char *vgpu_tunable(const char *name); /* getenv, but NULL when a control-plane config is mounted */
#define VP(p) vgpu_path((p), (char[512]){0}, 512)

/* boot.c */
extern volatile unsigned vgpu_fork_epoch; /* bumped in the child by a pthread_atfork handler */
void vgpu_boot(void);             /* == reference load_necessary_data (loader.c:2166) */
void vgpu_map_devices(void);      /* == reference init_devices_mapping (loader.c:2178) */
const char *vgpu_cu_err(CUresult r);
const char *vgpu_nv_err(nvmlReturn_t r);
void *vgpu_lookup_cuda_hook(const char *name, int want_ptsz);
void *vgpu_lookup_nvml_hook(const char *name);

/* config.c */
int vgpu_host_index_of_cuda(CUdevice dev);
int vgpu_nvml_index_of_cuda(CUdevice dev);
int vgpu_host_index_of_nvml(nvmlDevice_t dev);
nvmlDevice_t vgpu_nvml_handle_of_host(int host_index);
int vgpu_lock_gpu(int host_index);
void vgpu_unlock_gpu(int fd);
int vgpu_vmem_lock(int host_index, int write);
void vgpu_vmem_unlock(int fd, int host_index);
int vgpu_smutil_rdlock(int host_index);
void vgpu_smutil_unlock(int fd, int host_index);
/* device footprint of this library summed over the live processes of this container on GPU h;
 * publish != 0 (re)registers the calling process with publish_bytes first.  GPU lock must be held. */
uint64_t vgpu_self_registry(int h, uint64_t publish_bytes, int publish);
/* container membership flags for a list of device pids (VGPU_FLAG_*), per compatibility mode */
void vgpu_pid_flags(const uint32_t *pids, uint32_t n, uint8_t *flags);
/* same for the utilisation fold: client mode with an empty pids.config is not fatal there, the
 * reference just skips the samples (cuda_hook.c:1073); returns 0 in exactly that case */
int vgpu_pid_flags_util(const uint32_t *pids, uint32_t n, uint8_t *flags);

/* device.c - per-GPU device runtime (module, streams, pinned blocks) */
typedef struct vgpu_dev_rt {
  int host_index;
  int ready;   /* 1 once the module is loaded in `ctx`; -1 if bring-up failed (retried at retry_at) */
  unsigned fails;     /* consecutive failed bring-ups */
  uint64_t retry_at;  /* CLOCK_MONOTONIC second of the next attempt; 0 = never (deterministic failure) */
  CUdevice cuda_dev;
  CUcontext ctx;
  CUmodule mod;
  CUfunction k_clear, k_spill, k_copy_generic, k_quota, k_slab_insert, k_slab_remove, k_controller, k_sampler, k_gate, k_governor, k_refill, k_vslab;
  CUstream q_stream; /* quota / ledger kernels (app thread)     */
  CUstream s_stream; /* sampler + controller (tick thread); the governor in VGPU_B200_GOVERNOR=1 mode */
  CUstream p_stream; /* direct-API sampler runs; probe-only sampler beside the governor          */
  /* pinned, mapped blocks */
  vgpu_quota_req_t *q_req;  CUdeviceptr q_req_d;
  vgpu_quota_res_t *q_res;  CUdeviceptr q_res_d;
  vgpu_slab_res_t *slab_res; CUdeviceptr slab_res_d;
  vgpu_lim_host_t *lim_h;   CUdeviceptr lim_h_d;
  vgpu_util_req_t *u_req;   CUdeviceptr u_req_d;  /* utilisation publication (tick thread -> refill kernel) */
  vgpu_vslab_res_t *vs_res; CUdeviceptr vs_res_d; /* slab placement table answers */
  /* HBM */
  CUdeviceptr lim_d;  /* vgpu_lim_dev_t          */
  CUdeviceptr slab_d; /* vgpu_slab_slot_t[SLOTS] */
  CUdeviceptr vslab_d; /* vgpu_vslab_slot_t[VGPU_VSLAB_SLOTS]: placement table of the slab mode */
  struct vgpu_slab_host *vs_host; /* host half of the placement table (driver handles), slabmode.c */
  uint32_t vs_age;     /* allocation counter = age stamp */
  CUevent ev0, ev1;    /* timing of the spill / scrub kernels launched by the hooks */
  uint64_t self_bytes; /* measured device footprint of everything above */
  int q_req_self_set;  /* q_req->self_bytes was provided by the caller of vgpu_rt_quota */
  uint32_t seq;
  uint32_t q_longest;  /* longest list of the previous quota evaluation (sizes the armed launch) */
  int quota_armed;     /* allocation hooks launch the quota kernel ahead of the NVML queries */
  int gfx_valid;       /* q_req->graphics / gflags hold the graphics list of the previous evaluation */
  volatile long uva_live; /* records in the slab (host-side count; skip lookups when 0) */
  vgpu_slab_slot_t *uva_snap; /* records read back while the context is being torn down (vgpu_rt_context_before) */
  uint32_t uva_snap_n;
  pthread_mutex_t q_mu;
  int sm_num, max_thread_per_sm;
  int64_t total_cores;
  int memops64; /* cuStreamWaitValue64 usable */
  int ctx_is_primary; /* `ctx` is the device's primary context */
  uint32_t spill_chunk, spill_stages, spill_ctas_per_sm;
} vgpu_dev_rt;

vgpu_dev_rt *vgpu_rt_get(int host_index, CUdevice dev); /* bring up (needs a current ctx) */
vgpu_dev_rt *vgpu_rt_peek(int host_index);              /* NULL unless ready */
/* The tenant is about to destroy / reset / release a context: every runtime living in it stops
 * being used (its streams, module, HBM and pinned blocks die with the context).  Returns a
 * bitmask of the detached runtime slots for vgpu_rt_context_after(). */
unsigned vgpu_rt_context_before(CUcontext ctx, CUdevice dev, int primary);
/* still_alive != 0 (a primary-context release that left references): take the runtimes back */
void vgpu_rt_context_after(unsigned mask, int still_alive);
CUresult vgpu_rt_launch(vgpu_dev_rt *rt, CUfunction f, unsigned grid, unsigned block,
                        unsigned smem, CUstream s, void **params);

int vgpu_rt_quota(vgpu_dev_rt *rt, vgpu_quota_res_t *out);
uint32_t vgpu_rt_quota_arm(vgpu_dev_rt *rt);
void vgpu_rt_quota_publish(vgpu_dev_rt *rt, uint32_t seq);
int vgpu_rt_quota_collect(vgpu_dev_rt *rt, uint32_t seq, vgpu_quota_res_t *out);
int vgpu_rt_slab_insert(vgpu_dev_rt *rt, CUdeviceptr dptr, uint64_t bytes);
int vgpu_rt_slab_remove(vgpu_dev_rt *rt, CUdeviceptr dptr, uint64_t *bytes);
int vgpu_stale_uva_remove(CUdeviceptr dptr, uint64_t *bytes); /* records that outlived their context (device.c) */
CUresult vgpu_rt_clear(vgpu_dev_rt *rt, CUdeviceptr dst, size_t bytes, CUstream s);
CUresult vgpu_rt_spill(vgpu_dev_rt *rt, CUdeviceptr dst, CUdeviceptr src, size_t bytes, CUstream s);

/* slabmode.c - VGPU_B200_SLAB=1: cuMemAlloc of an oversold device served from VMM-backed slabs */
int vgpu_slab_mode(void);
/* path = the quota kernel's decision (VGPU_PATH_GPU / _UVA).  CUDA_SUCCESS: *dptr is a slab;
 * CUDA_ERROR_NOT_SUPPORTED: slab mode cannot serve this request, use the plain path. */
CUresult vgpu_slab_alloc(vgpu_dev_rt *rt, CUdevice dev, int path, CUdeviceptr *dptr, size_t bytes, int *recorded_uva);
/* 1 if dptr is a slab (then *out is the result of freeing it and *was_uva / *bytes describe its
 * accounting), 0 if it is not one */
int vgpu_slab_free(vgpu_dev_rt *rt, CUdeviceptr dptr, CUresult *out, int *was_uva, uint64_t *bytes);
void vgpu_slab_forget(vgpu_dev_rt *rt); /* the context died: drop the host half */

/* limiter.c */
void vgpu_limiter_start(void); /* == reference initialization() (cuda_hook.c:566) */
void vgpu_limiter_detach(int host_index); /* tick + watchdog threads stop touching this device's runtime */
void vgpu_limiter_attach(int host_index, int forget_streams);
void vgpu_limiter_before_blocking_call(vgpu_dev_rt *rt); /* wait in user space while tenant work is parked behind the gate */
void vgpu_limiter_quiesce(vgpu_dev_rt *rt); /* device-wide sync ahead: governor retires once nothing is parked */
void vgpu_limiter_resume(vgpu_dev_rt *rt, int everything_completed); /* the sync returned */

/* metrics.c */
enum { VM_RATE_GATED, VM_RATE_FAST, VM_OOM_LIMIT, VM_OOM_DRIVER, VM_UVA_FALLBACK, VM_LOCK_TIMEOUT,
       VM_QUOTA_KERNELS, VM_SAMPLER_LAUNCHES, VM_SCRUBBED_BYTES, VM_WATCHDOG_LOANS, VM_SAMPLER_SKIPPED,
       VM_SPILL_BYTES, VM_SPILL_NS, VM_SCRUB_NS, VM_PROMOTE_BYTES, VM_PROMOTE_NS, VM_SLAB_ALLOCS, VM_SLAB_DEMOTIONS, VM_COUNT };
void vgpu_metric_add(int host_index, int which, uint64_t v);
uint64_t vgpu_metric_get(int host_index, int which);

#endif
