/*
 * direct.c - a tenant that is LINKED against libcuda.so.1 / libnvidia-ml.so.1 (no dlopen, no
 * dlsym, no cuGetProcAddress): interception must then come from plain symbol interposition of
 * the preloaded library.  TEST INFRASTRUCTURE.
 */
#include <stddef.h>
#include <stdio.h>
typedef int CUresult;
typedef unsigned long long CUdeviceptr;
extern CUresult cuInit(unsigned);
extern CUresult cuDeviceGet(int *, int);
extern CUresult cuDevicePrimaryCtxRetain(void **, int);
extern CUresult cuCtxSetCurrent(void *);
extern CUresult cuDeviceTotalMem_v2(size_t *, int);
extern CUresult cuMemGetInfo_v2(size_t *, size_t *);
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuLaunchKernel(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **);
extern int nvmlInit_v2(void);
extern int nvmlDeviceGetHandleByIndex_v2(unsigned, void **);
typedef struct { unsigned long long total, free, used; } nvmem_t;
extern int nvmlDeviceGetMemoryInfo(void *, nvmem_t *);

int main(void) {
  int dev = 0, r1, r2, r3, r4;
  void *ctx = NULL, *nv = NULL;
  size_t tot = 0, fr = 0;
  CUdeviceptr p = 0, q = 0;
  r1 = cuInit(0);
  r2 = cuDeviceGet(&dev, 0);
  r3 = cuDevicePrimaryCtxRetain(&ctx, dev);
  r4 = cuCtxSetCurrent(ctx);
  printf("init %d %d %d %d\n", r1, r2, r3, r4);
  r1 = nvmlInit_v2();
  r2 = nvmlDeviceGetHandleByIndex_v2(0, &nv);
  printf("nvml %d %d\n", r1, r2);
  r1 = cuDeviceTotalMem_v2(&tot, dev);
  printf("totalmem %d %zu\n", r1, tot);
  r1 = cuMemAlloc_v2(&p, 600u << 20);
  printf("alloc %d\n", r1);
  r1 = cuMemAlloc_v2(&q, 600u << 20);
  printf("alloc %d\n", r1);
  r1 = cuMemGetInfo_v2(&fr, &tot);
  printf("meminfo %d %zu %zu\n", r1, fr, tot);
  nvmem_t m = {0, 0, 0};
  r1 = nvmlDeviceGetMemoryInfo(nv, &m);
  printf("nvmlinfo %d %llu %llu %llu\n", r1, m.total, m.free, m.used);
  int ok = 0;
  for (int i = 0; i < 1000; i++) ok += cuLaunchKernel(NULL, 1, 1, 1, 1, 1, 1, 0, NULL, NULL, NULL) == 0;
  printf("launch %d\n", ok);
  r1 = cuMemFree_v2(p);
  printf("free %d\n", r1);
  return 0;
}
