/*
 * vmm_probe.c - is a VMM-backed slab that can be re-pointed from HBM to host memory under the same
 * virtual address feasible on this box?  (round-2 design probe for the VGPU_B200_SLAB mode)
 *   gcc -O2 -I/usr/local/cuda/include -o vmm_probe vmm_probe.c -ldl && ./vmm_probe
 * Prints one JSON object with per-call timings (us) and the outcome of each step.
 */
#define _GNU_SOURCE
#include <cuda.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
#define SYM(name) __typeof__(name) *p_##name = (__typeof__(name) *)dlsym(h, #name); if (!p_##name) { printf("{\"error\": \"missing %s\"}\n", #name); return 1; }
#define T(var, call) do { double t0_ = now_us(); CUresult r_ = (call); var = now_us() - t0_; if (r_ != CUDA_SUCCESS) { printf("{\"error\": \"%s -> %d\"}\n", #call, (int)r_); return 1; } } while (0)
#define TRY(rc, var, call) do { double t0_ = now_us(); rc = (call); var = now_us() - t0_; } while (0)

int main(int argc, char **argv) {
  size_t bytes = argc > 1 ? (size_t)atol(argv[1]) << 20 : (size_t)64 << 20;
  void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { printf("{\"error\": \"no libcuda\"}\n"); return 1; }
  SYM(cuInit) SYM(cuDeviceGet) SYM(cuCtxCreate_v2) SYM(cuMemGetAllocationGranularity) SYM(cuMemAddressReserve)
  SYM(cuMemCreate) SYM(cuMemMap) SYM(cuMemSetAccess) SYM(cuMemUnmap) SYM(cuMemRelease) SYM(cuMemAddressFree)
  SYM(cuMemsetD8_v2) SYM(cuMemcpyDtoH_v2) SYM(cuMemcpyDtoD_v2) SYM(cuCtxSynchronize) SYM(cuMemAlloc_v2) SYM(cuMemFree_v2)
  SYM(cuMemGetInfo_v2)
  CUdevice dev; CUcontext ctx; double t;
  T(t, p_cuInit(0));
  T(t, p_cuDeviceGet(&dev, 0));
  T(t, p_cuCtxCreate_v2(&ctx, 0, dev));
  CUmemAllocationProp dprop; memset(&dprop, 0, sizeof dprop);
  dprop.type = CU_MEM_ALLOCATION_TYPE_PINNED; dprop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; dprop.location.id = dev;
  size_t gran = 0, gran_rec = 0;
  T(t, p_cuMemGetAllocationGranularity(&gran, &dprop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  T(t, p_cuMemGetAllocationGranularity(&gran_rec, &dprop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  CUmemAllocationProp hprop; memset(&hprop, 0, sizeof hprop);
  hprop.type = CU_MEM_ALLOCATION_TYPE_PINNED; hprop.location.type = CU_MEM_LOCATION_TYPE_HOST_NUMA; hprop.location.id = 0;
  size_t hgran = 0; CUresult rc; double t_hgran;
  TRY(rc, t_hgran, p_cuMemGetAllocationGranularity(&hgran, &hprop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  int host_numa_ok = rc == CUDA_SUCCESS;
  size_t free0, total0, free1;
  p_cuMemGetInfo_v2(&free0, &total0);
  double t_reserve, t_create, t_map, t_access, t_memset, t_hcreate = 0, t_hmap = 0, t_haccess = 0, t_copy = 0, t_unmap, t_remap = 0, t_raccess = 0, t_release, t_malloc, t_mfree;
  CUdeviceptr va = 0, stage_va = 0, plain = 0;
  CUmemGenericAllocationHandle dh = 0, hh = 0;
  T(t_malloc, p_cuMemAlloc_v2(&plain, bytes));
  T(t_mfree, p_cuMemFree_v2(plain));
  T(t_reserve, p_cuMemAddressReserve(&va, bytes, 0, 0, 0));
  T(t_create, p_cuMemCreate(&dh, bytes, &dprop, 0));
  T(t_map, p_cuMemMap(va, bytes, 0, dh, 0));
  CUmemAccessDesc acc; memset(&acc, 0, sizeof acc);
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = dev; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  T(t_access, p_cuMemSetAccess(va, bytes, &acc, 1));
  T(t_memset, p_cuMemsetD8_v2(va, 0x5A, bytes));
  p_cuCtxSynchronize();
  p_cuMemGetInfo_v2(&free1, &total0);
  int host_map_ok = 0, remap_ok = 0, data_ok = 0;
  if (host_numa_ok) {
    TRY(rc, t_hcreate, p_cuMemCreate(&hh, bytes, &hprop, 0));
    if (rc == CUDA_SUCCESS) {
      T(t, p_cuMemAddressReserve(&stage_va, bytes, 0, 0, 0));
      TRY(rc, t_hmap, p_cuMemMap(stage_va, bytes, 0, hh, 0));
      if (rc == CUDA_SUCCESS) {
        TRY(rc, t_haccess, p_cuMemSetAccess(stage_va, bytes, &acc, 1));
        host_map_ok = rc == CUDA_SUCCESS;
      }
    } else host_numa_ok = -(int)rc;
  }
  if (host_map_ok) {
    /* demote: copy HBM slab -> host-backed staging mapping, swap the backing under `va` */
    T(t_copy, p_cuMemcpyDtoD_v2(stage_va, va, bytes));
    p_cuCtxSynchronize();
    T(t_unmap, p_cuMemUnmap(va, bytes));
    T(t, p_cuMemUnmap(stage_va, bytes));
    TRY(rc, t_remap, p_cuMemMap(va, bytes, 0, hh, 0));
    if (rc == CUDA_SUCCESS) {
      TRY(rc, t_raccess, p_cuMemSetAccess(va, bytes, &acc, 1));
      remap_ok = rc == CUDA_SUCCESS;
    }
    if (remap_ok) {
      unsigned char probe[64];
      T(t, p_cuMemcpyDtoH_v2(probe, va + bytes - 64, 64));
      data_ok = 1;
      for (int i = 0; i < 64; i++) data_ok &= probe[i] == 0x5A;
      /* the device must be able to write through the host-backed mapping too */
      T(t, p_cuMemsetD8_v2(va, 0x33, 4096));
      p_cuCtxSynchronize();
      T(t, p_cuMemcpyDtoH_v2(probe, va, 64));
      data_ok &= probe[0] == 0x33;
    }
  } else {
    T(t_unmap, p_cuMemUnmap(va, bytes));
  }
  size_t free2 = 0;
  T(t_release, p_cuMemRelease(dh));
  p_cuMemGetInfo_v2(&free2, &total0);
  printf("{\"bytes\": %zu, \"granularity_min\": %zu, \"granularity_recommended\": %zu, \"host_numa_granularity\": %zu, "
         "\"host_numa_ok\": %d, \"host_map_ok\": %d, \"remap_same_va_ok\": %d, \"data_ok\": %d, "
         "\"free_before\": %zu, \"free_with_device_slab\": %zu, \"free_after_release\": %zu, "
         "\"us\": {\"cuMemAlloc\": %.1f, \"cuMemFree\": %.1f, \"reserve\": %.1f, \"create_device\": %.1f, \"map\": %.1f, \"set_access\": %.1f, "
         "\"memset\": %.1f, \"create_host\": %.1f, \"map_host\": %.1f, \"set_access_host\": %.1f, \"copy_d2h_mapped\": %.1f, "
         "\"unmap\": %.1f, \"remap\": %.1f, \"remap_access\": %.1f, \"release\": %.1f}}\n",
         bytes, gran, gran_rec, hgran, host_numa_ok, host_map_ok, remap_ok, data_ok, free0, free1, free2, t_malloc, t_mfree, t_reserve,
         t_create, t_map, t_access, t_memset, t_hcreate, t_hmap, t_haccess, t_copy, t_unmap, t_remap, t_raccess, t_release);
  return 0;
}
