/*
 * mtstorm.c - multi-threaded allocation / launch mix for the interception library.
 * TEST INFRASTRUCTURE.
 *
 * T threads share one context; each runs K steps of a seeded random mix of cuMemAlloc,
 * cuMemAllocManaged, cuMemFree, cuMemGetInfo, nvmlDeviceGetMemoryInfo and cuLaunchKernel.  Every
 * thread frees what it still holds at the end, so the final figures are interleaving-independent
 * and can be compared between two libraries; per-call results may only be SUCCESS or OUT_OF_MEMORY.
 *   mtstorm [--threads T] [--steps K] [--seed S]
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef struct { unsigned long long total, free, used; } nvmem_t;

static CUresult (*p_alloc)(CUdeviceptr *, size_t);
static CUresult (*p_managed)(CUdeviceptr *, size_t, unsigned);
static CUresult (*p_free)(CUdeviceptr);
static CUresult (*p_meminfo)(size_t *, size_t *);
static CUresult (*p_set)(void *);
static CUresult (*p_launch)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **);
static int (*n_meminfo)(void *, nvmem_t *);
static void *g_ctx, *g_nvdev;
static int g_steps = 400;
static unsigned g_seed = 1;

typedef struct { int id; long ok, oom, unexpected, bad_report; } worker_t;

static void *worker(void *arg) {
  worker_t *w = (worker_t *)arg;
  uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(w->id + 1) ^ g_seed;
  CUdeviceptr held[32];
  int nheld = 0;
  p_set(g_ctx);
  for (int i = 0; i < g_steps; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    unsigned k = (unsigned)(s % 100);
    size_t bytes = (size_t)(((s >> 8) % 48) + 1) << 20; /* 1..48 MiB */
    if (k < 40 && nheld < 32) {
      CUdeviceptr p = 0;
      CUresult r = (k < 30) ? p_alloc(&p, bytes) : p_managed(&p, bytes, 1);
      if (r == 0) { held[nheld++] = p; w->ok++; }
      else if (r == 2) w->oom++;
      else w->unexpected++;
    } else if (k < 75 && nheld > 0) {
      int j = (int)((s >> 20) % (unsigned)nheld);
      CUresult r = p_free(held[j]);
      if (r != 0) w->unexpected++;
      held[j] = held[--nheld];
    } else if (k < 85) {
      size_t fr = 0, tot = 0;
      if (p_meminfo(&fr, &tot) != 0 || fr > tot) w->bad_report++;
    } else if (k < 92) {
      nvmem_t m;
      if (n_meminfo(g_nvdev, &m) != 0 || m.used > m.total || m.free + m.used != m.total) w->bad_report++;
    } else {
      p_launch(NULL, 4, 1, 1, 1, 1, 1, 0, NULL, NULL, NULL);
    }
  }
  while (nheld > 0)
    if (p_free(held[--nheld]) != 0) w->unexpected++;
  return NULL;
}

int main(int argc, char **argv) {
  int threads = 8;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--steps") && i + 1 < argc) g_steps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--seed") && i + 1 < argc) g_seed = (unsigned)atoi(argv[++i]);
  }
  if (threads < 1 || threads > 64) threads = 8;
  void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL), *hn = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h || !hn) { fprintf(stderr, "mtstorm: %s\n", dlerror()); return 2; }
  CUresult (*p_init)(unsigned) = dlsym(h, "cuInit");
  CUresult (*p_get)(int *, int) = dlsym(h, "cuDeviceGet");
  CUresult (*p_retain)(void **, int) = dlsym(h, "cuDevicePrimaryCtxRetain");
  int (*n_init)(void) = dlsym(hn, "nvmlInit_v2");
  int (*n_h)(unsigned, void **) = dlsym(hn, "nvmlDeviceGetHandleByIndex_v2");
  p_set = dlsym(h, "cuCtxSetCurrent");
  p_alloc = dlsym(h, "cuMemAlloc_v2");
  p_managed = dlsym(h, "cuMemAllocManaged");
  p_free = dlsym(h, "cuMemFree_v2");
  p_meminfo = dlsym(h, "cuMemGetInfo_v2");
  p_launch = dlsym(h, "cuLaunchKernel");
  n_meminfo = dlsym(hn, "nvmlDeviceGetMemoryInfo");
  int dev = 0;
  if (p_init(0) || p_get(&dev, 0) || p_retain(&g_ctx, dev) || p_set(g_ctx) || n_init() || n_h(0, &g_nvdev)) {
    fprintf(stderr, "mtstorm: bring-up failed\n");
    return 2;
  }
  pthread_t tid[64];
  worker_t w[64];
  memset(w, 0, sizeof w);
  for (int t = 0; t < threads; t++) { w[t].id = t; pthread_create(&tid[t], NULL, worker, &w[t]); }
  long ok = 0, oom = 0, unexpected = 0, bad = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(tid[t], NULL);
    ok += w[t].ok; oom += w[t].oom; unexpected += w[t].unexpected; bad += w[t].bad_report;
  }
  size_t fr = 0, tot = 0;
  nvmem_t m = {0, 0, 0};
  CUresult r1 = p_meminfo(&fr, &tot);
  int r2 = n_meminfo(g_nvdev, &m);
  /* interleaving-dependent counters go to stderr, the comparable final state to stdout */
  fprintf(stderr, "mtstorm: ok %ld oom %ld\n", ok, oom);
  printf("final meminfo %d free %zu total %zu nvml %d used %llu total %llu unexpected %ld bad_reports %ld\n", r1, fr, tot, r2,
         m.used, m.total, unexpected, bad);
  return (unexpected || bad) ? 1 : 0;
}
