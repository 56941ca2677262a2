/*
 * allocstorm.c - allocator storm (SURVEY.md 8d "Reference CPU timing"): N x { cuMemAlloc(bytes),
 * cuMemFree } through the dynamic linker, per-call latency percentiles, JSON on stdout.
 * TEST / BENCH INFRASTRUCTURE.
 *   allocstorm [--n N] [--bytes B] [--device D]
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
static inline uint64_t now_ns(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return (x > y) - (x < y);
}
int main(int argc, char **argv) {
  long n = 2000, bytes = 1 << 20;
  int device = 0;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--n") && i + 1 < argc) n = atol(argv[++i]);
    else if (!strcmp(argv[i], "--bytes") && i + 1 < argc) bytes = atol(argv[++i]);
    else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
  }
  void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "allocstorm: %s\n", dlerror()); return 2; }
  CUresult (*p_init)(unsigned) = dlsym(h, "cuInit");
  CUresult (*p_get)(int *, int) = dlsym(h, "cuDeviceGet");
  CUresult (*p_retain)(void **, int) = dlsym(h, "cuDevicePrimaryCtxRetain");
  CUresult (*p_set)(void *) = dlsym(h, "cuCtxSetCurrent");
  CUresult (*p_alloc)(CUdeviceptr *, size_t) = dlsym(h, "cuMemAlloc_v2");
  CUresult (*p_free)(CUdeviceptr) = dlsym(h, "cuMemFree_v2");
  int dev = 0;
  void *ctx = NULL;
  if (p_init(0) || p_get(&dev, device) || p_retain(&ctx, dev) || p_set(ctx)) { fprintf(stderr, "allocstorm: init failed\n"); return 3; }
  CUdeviceptr p = 0;
  for (int i = 0; i < 20; i++) { if (p_alloc(&p, (size_t)bytes) == 0) p_free(p); } /* warm-up + bring-up */
  uint32_t *la = malloc(sizeof(uint32_t) * (size_t)n), *lf = malloc(sizeof(uint32_t) * (size_t)n);
  long fails = 0;
  uint64_t t0 = now_ns();
  for (long i = 0; i < n; i++) {
    uint64_t a = now_ns();
    CUresult r = p_alloc(&p, (size_t)bytes);
    uint64_t b = now_ns();
    if (r == 0) p_free(p); else fails++;
    uint64_t c = now_ns();
    la[i] = (uint32_t)(b - a);
    lf[i] = (uint32_t)(c - b);
  }
  uint64_t t1 = now_ns();
  qsort(la, (size_t)n, sizeof(uint32_t), cmp_u32);
  qsort(lf, (size_t)n, sizeof(uint32_t), cmp_u32);
  printf("{\"pairs\": %ld, \"bytes\": %ld, \"pairs_per_s\": %.1f, \"alloc_p50_ns\": %u, \"alloc_p99_ns\": %u, "
         "\"free_p50_ns\": %u, \"free_p99_ns\": %u, \"fails\": %ld}\n",
         n, bytes, n / ((t1 - t0) * 1e-9), la[n / 2], la[(long)(n * 0.99)], lf[n / 2], lf[(long)(n * 0.99)], fails);
  return 0;
}
