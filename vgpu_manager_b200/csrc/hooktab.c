/*
 * hooktab.c - the intercepted symbol surface as a name -> entry-point table.
 *
 * This is the drop-in boundary of SURVEY.md 8(b): the 37 CUDA names of reference
 * library/src/cuda_hook.c:243-281 and the 7 NVML names of library/src/nvml_hook.c:20-28, plus
 * cuCtxSynchronize (needed so a resident sampler kernel never delays a tenant's device
 * synchronisation), the other blocking calls a throttled tenant commonly makes (cuStreamSynchronize,
 * cuEventSynchronize, synchronous cuMemcpyDtoH: they wait in user space while work is parked behind
 * the gate, limiter.c), cuStreamDestroy and the context teardown entry points (the device-resident
 * state lives in the tenant's context) and nvmlDeviceGetUtilizationRates (named by BASELINE.json; pure forward in
 * the reference, library/src/nvml_originals.c:698-702).
 * The entry points are only address-taken here, hence the untyped declarations.
 */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#define HOOK_DECL(n) extern void n(void);
#define VGPU_CUDA_HOOKS(X)                                                                    \
  X(cuDriverGetVersion) X(cuInit) X(cuGetProcAddress) X(cuGetProcAddress_v2)                  \
  X(cuMemAllocManaged) X(cuMemAlloc_v2) X(cuMemAlloc) X(cuMemAllocPitch_v2) X(cuMemAllocPitch) \
  X(cuArrayCreate_v2) X(cuArrayCreate) X(cuArray3DCreate_v2) X(cuArray3DCreate)               \
  X(cuMipmappedArrayCreate) X(cuDeviceTotalMem_v2) X(cuDeviceTotalMem) X(cuMemGetInfo_v2)     \
  X(cuMemGetInfo) X(cuLaunchKernel_ptsz) X(cuLaunchKernel) X(cuLaunchKernelEx_ptsz)           \
  X(cuLaunchKernelEx) X(cuLaunch) X(cuLaunchCooperativeKernel_ptsz) X(cuLaunchCooperativeKernel) \
  X(cuLaunchGrid) X(cuLaunchGridAsync) X(cuFuncSetBlockShape) X(cuMemAllocAsync)              \
  X(cuMemAllocAsync_ptsz) X(cuMemCreate) X(cuMemAllocFromPoolAsync)                           \
  X(cuMemAllocFromPoolAsync_ptsz) X(cuMemFree_v2) X(cuMemFree) X(cuMemFreeAsync)              \
  X(cuMemFreeAsync_ptsz) X(cuCtxSynchronize) X(cuStreamDestroy_v2)                             \
  X(cuStreamSynchronize) X(cuStreamSynchronize_ptsz) X(cuEventSynchronize)                      \
  X(cuMemcpyDtoH_v2) X(cuMemcpyDtoH_v2_ptds) X(cuMemcpyHtoD_v2) X(cuMemcpyHtoD_v2_ptds)         \
  X(cuMemcpy) X(cuMemcpy_ptds)                                                                  \
  X(cuCtxDestroy_v2) X(cuCtxDestroy) X(cuDevicePrimaryCtxReset_v2) X(cuDevicePrimaryCtxReset)  \
  X(cuDevicePrimaryCtxRelease_v2) X(cuDevicePrimaryCtxRelease)
/* opt-in (VGPU_B200_GRAPH_LIMIT=1): only then are these names substituted in dlsym /
 * cuGetProcAddress answers; the reference forwards all of them (cuda_originals.c:2953-3040) */
#define VGPU_GRAPH_HOOKS(X)                                                                   \
  X(cuGraphInstantiateWithFlags) X(cuGraphInstantiateWithParams) X(cuGraphInstantiateWithParams_ptsz) \
  X(cuGraphLaunch) X(cuGraphLaunch_ptsz) X(cuGraphExecDestroy)
#define VGPU_NVML_HOOKS(X)                                                                    \
  X(nvmlInit) X(nvmlInit_v2) X(nvmlInitWithFlags) X(nvmlDeviceGetMemoryInfo)                  \
  X(nvmlDeviceGetMemoryInfo_v2) X(nvmlDeviceSetComputeMode) X(nvmlDeviceGetPersistenceMode)   \
  X(nvmlDeviceGetUtilizationRates)
VGPU_CUDA_HOOKS(HOOK_DECL)
VGPU_GRAPH_HOOKS(HOOK_DECL)
VGPU_NVML_HOOKS(HOOK_DECL)
extern int vgpu_graph_limit_enabled(void);

typedef struct { const char *name; void *fn; } hook_ent;
#define HOOK_ENT(n) {#n, (void *)n},
static const hook_ent g_cuda_hooks[] = {VGPU_CUDA_HOOKS(HOOK_ENT)};
static const hook_ent g_graph_hooks[] = {VGPU_GRAPH_HOOKS(HOOK_ENT)};
static const hook_ent g_nvml_hooks[] = {VGPU_NVML_HOOKS(HOOK_ENT)};

static void *find_hook(const hook_ent *t, size_t n, const char *name) {
  for (size_t i = 0; i < n; i++)
    if (!strcmp(t[i].name, name)) return t[i].fn;
  return NULL;
}

void *vgpu_lookup_cuda_hook(const char *name, int want_ptsz) {
  size_t n = sizeof g_cuda_hooks / sizeof g_cuda_hooks[0];
  /* cuGetProcAddress callers ask for the base name; cuStreamDestroy has had the _v2 ABI since
   * CUDA 4.0 and the driver hands out exactly that for it */
  if (!strcmp(name, "cuStreamDestroy")) name = "cuStreamDestroy_v2";
  if (!strcmp(name, "cuCtxDestroy")) name = "cuCtxDestroy_v2";                           /* v2 ABI since CUDA 4.0 */
  /* the synchronous copies: v2 ABI since CUDA 3.2, per-thread-default-stream suffix is _ptds */
  if (!strcmp(name, "cuMemcpyDtoH")) name = want_ptsz ? "cuMemcpyDtoH_v2_ptds" : "cuMemcpyDtoH_v2";
  else if (!strcmp(name, "cuMemcpyHtoD")) name = want_ptsz ? "cuMemcpyHtoD_v2_ptds" : "cuMemcpyHtoD_v2";
  else if (!strcmp(name, "cuMemcpy")) name = want_ptsz ? "cuMemcpy_ptds" : "cuMemcpy";
  if (!strcmp(name, "cuDevicePrimaryCtxReset")) name = "cuDevicePrimaryCtxReset_v2";     /* since CUDA 11.0 */
  if (!strcmp(name, "cuDevicePrimaryCtxRelease")) name = "cuDevicePrimaryCtxRelease_v2"; /* since CUDA 11.0 */
  if (!strncmp(name, "cuGraph", 7)) {
    if (!vgpu_graph_limit_enabled()) return NULL;
    /* since CUDA 12.0 the base name is the WithFlags entry point (cuda.h maps it); the legacy
     * 5-argument ABI of older toolkits is left to the driver */
    if (!strcmp(name, "cuGraphInstantiate")) name = "cuGraphInstantiateWithFlags";
    size_t ng = sizeof g_graph_hooks / sizeof g_graph_hooks[0];
    if (want_ptsz) {
      char alt[96];
      snprintf(alt, sizeof alt, "%s_ptsz", name);
      void *f = find_hook(g_graph_hooks, ng, alt);
      if (f) return f;
    }
    return find_hook(g_graph_hooks, ng, name);
  }
  if (want_ptsz) {
    char alt[96];
    snprintf(alt, sizeof alt, "%s_ptsz", name);
    void *f = find_hook(g_cuda_hooks, n, alt);
    if (f) return f;
  }
  return find_hook(g_cuda_hooks, n, name);
}

void *vgpu_lookup_nvml_hook(const char *name) {
  return find_hook(g_nvml_hooks, sizeof g_nvml_hooks / sizeof g_nvml_hooks[0], name);
}

