/*
 * lifecycle.c - hooks on the entry points that end a CUDA context.
 *
 * No reference counterpart: library/src keeps no device state, so cuCtxDestroy and friends are
 * plain forwards there (cuda_originals.c).  Here the token bucket, the UVA slab, three streams, the
 * module and four pinned blocks live in the tenant's context (csrc/device.c), so the runtime has to
 * be detached before the driver frees them and rebuilt by the next hooked call.
 */
#include "vgpu_internal.h"

static CUresult ctx_gone(CUresult (*real)(CUcontext), CUcontext ctx) {
  vgpu_boot();
  if (unlikely(!real)) return CUDA_ERROR_NOT_FOUND;
  unsigned mask = vgpu_rt_context_before(ctx, 0, 0);
  CUresult r = real(ctx);
  vgpu_rt_context_after(mask, r != CUDA_SUCCESS);
  return r;
}
static CUresult primary_gone(CUresult (*real)(CUdevice), CUdevice dev, int release) {
  vgpu_boot();
  if (unlikely(!real)) return CUDA_ERROR_NOT_FOUND;
  unsigned mask = vgpu_rt_context_before(NULL, dev, 1);
  CUresult r = real(dev);
  int alive = (r != CUDA_SUCCESS);
  if (!alive && release && mask) { /* a release only destroys the context when it was the last reference */
    unsigned int fl = 0;
    int active = 0;
    if (R.cuDevicePrimaryCtxGetState && R.cuDevicePrimaryCtxGetState(dev, &fl, &active) == CUDA_SUCCESS) alive = active;
  }
  vgpu_rt_context_after(mask, alive);
  return r;
}
VGPU_EXPORT CUresult cuCtxDestroy_v2(CUcontext ctx) { return ctx_gone(R.cuCtxDestroy_v2 ? R.cuCtxDestroy_v2 : R.cuCtxDestroy, ctx); }
VGPU_EXPORT CUresult cuCtxDestroy(CUcontext ctx) { return ctx_gone(R.cuCtxDestroy ? R.cuCtxDestroy : R.cuCtxDestroy_v2, ctx); }
VGPU_EXPORT CUresult cuDevicePrimaryCtxReset_v2(CUdevice dev) {
  return primary_gone(R.cuDevicePrimaryCtxReset_v2 ? R.cuDevicePrimaryCtxReset_v2 : R.cuDevicePrimaryCtxReset, dev, 0);
}
VGPU_EXPORT CUresult cuDevicePrimaryCtxReset(CUdevice dev) {
  return primary_gone(R.cuDevicePrimaryCtxReset ? R.cuDevicePrimaryCtxReset : R.cuDevicePrimaryCtxReset_v2, dev, 0);
}
VGPU_EXPORT CUresult cuDevicePrimaryCtxRelease_v2(CUdevice dev) {
  return primary_gone(R.cuDevicePrimaryCtxRelease_v2 ? R.cuDevicePrimaryCtxRelease_v2 : R.cuDevicePrimaryCtxRelease, dev, 1);
}
VGPU_EXPORT CUresult cuDevicePrimaryCtxRelease(CUdevice dev) {
  return primary_gone(R.cuDevicePrimaryCtxRelease ? R.cuDevicePrimaryCtxRelease : R.cuDevicePrimaryCtxRelease_v2, dev, 1);
}
