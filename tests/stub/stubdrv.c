/*
 * stubdrv.c - a fake NVIDIA driver (libcuda.so.1 + libnvidia-ml.so.1 in one object) for
 * CPU-only plumbing tests.  TEST INFRASTRUCTURE, never shipped.
 *
 * The reference has no mock driver; this one exists so that both the
 * reference library (oracle/_ref) and the B200 library can be driven through the identical
 * scripted scenario on a box without a GPU (BASELINE.json configs[0]) and their outputs diffed.
 *
 *  - "device memory" is host memory (mmap, lazily committed), so device pointers are valid
 *    host pointers and pinned/mapped host blocks map to themselves;
 *  - per-process usedGpuMemory = STUB_CTX_BYTES + sum of live device allocations of this
 *    process; extra tenants come from STUB_OTHER_PROCS="pid:bytes:c|g|cg,...";
 *  - SM utilisation model: STUB_UTIL="fixed:N" or "closed:K" (util% = launches in the last
 *    second * K / 1000, capped at 100) - a fixed value above the cap deadlocks the
 *    reference's storm, hence the closed-loop default;
 *  - the fake GPU can "run" the B200 library's kernels: cuModuleGetFunction() resolves the
 *    entry names of vgpu_manager_b200/csrc/kernel_abi.h and cuLaunchKernel() executes them
 *    with the CPU oracle (oracle/vgpu_oracle.c).  That makes the host logic testable here; the
 *    real kernels are tested against the same oracle on a B200 (tests/test_gpu_*.py).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include "../../oracle/vgpu_oracle.h"
#include "../../vgpu_manager_b200/csrc/kernel_abi.h"

#define EXPORT __attribute__((visibility("default")))
typedef int CUresult;
typedef int CUdevice;
typedef unsigned long long CUdeviceptr;
typedef int nvmlReturn_t;

/* ------------------------------------------------------------------ state */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static int g_inited, g_gpu_count = 1;
static uint64_t g_total_mem = 180ull << 30, g_phys_mem, g_ctx_bytes;
typedef struct { void *p; size_t n; int kind; int dev; } alloc_t; /* kind 0 device, 1 managed, 2 host, 3 vmm handle */
static alloc_t *g_allocs;
static size_t g_nallocs, g_callocs;
static uint64_t g_dev_bytes[16];
static __thread int t_cur_dev = 0;
static volatile int g_parked; /* callers blocked behind the gate (the fake GPU's "parked streams") */
static void *volatile g_parked_streams[64]; /* which streams those are (cuStreamQuery answers NOT_READY for them) */
static __thread int t_has_ctx = 0;
static int g_any_ctx;            /* this process has (had) a context somewhere: it shows up in NVML's lists */
static int g_ctx_on[16];         /* ... on this device */

typedef struct { uint32_t pid; uint64_t bytes; int compute, graphics; int sm; } other_t;
static other_t g_others[64];
static int g_nothers;

static int g_util_mode = 1;        /* 0 fixed, 1 closed loop */
static int g_util_fixed = 0;
static double g_util_k = 0.05;     /* percent per launch/second/1000 -> see below */
#define NB 128
static uint64_t g_bucket_ms[NB];
static uint32_t g_bucket_cnt[NB];
static volatile uint64_t g_launches;

static uint64_t now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000 + ts.tv_nsec / 1000000;
}
static uint64_t now_us_wall(void) {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return (uint64_t)tv.tv_sec * 1000000ull + tv.tv_usec;
}

static void note_launch(void) {
  uint64_t ms = now_ms() / 10; /* 10 ms buckets */
  uint32_t i = (uint32_t)(ms % NB);
  if (g_bucket_ms[i] != ms) { g_bucket_ms[i] = ms; g_bucket_cnt[i] = 0; }
  g_bucket_cnt[i]++;
  __sync_fetch_and_add(&g_launches, 1);
}

static int current_util(void) {
  if (g_util_mode == 0) return g_util_fixed;
  uint64_t ms = now_ms() / 10, n = 0;
  for (int i = 0; i < NB; i++)
    if (g_bucket_ms[i] + 100 > ms) n += g_bucket_cnt[i];
  double u = (double)n * g_util_k / 1000.0;
  return u > 100 ? 100 : (int)u;
}

static void stub_init(void) {
  if (g_inited) return;
  pthread_mutex_lock(&g_mu);
  if (!g_inited) {
    const char *s;
    if ((s = getenv("STUB_GPU_COUNT"))) g_gpu_count = atoi(s);
    if (g_gpu_count < 1) g_gpu_count = 1;
    if (g_gpu_count > 16) g_gpu_count = 16;
    if ((s = getenv("STUB_TOTAL_MEM"))) g_total_mem = strtoull(s, NULL, 10);
    g_phys_mem = g_total_mem;
    if ((s = getenv("STUB_PHYS_MEM"))) g_phys_mem = strtoull(s, NULL, 10);
    if ((s = getenv("STUB_CTX_BYTES"))) g_ctx_bytes = strtoull(s, NULL, 10);
    if ((s = getenv("STUB_OTHER_PROCS"))) {
      char *dup = strdup(s), *save = NULL;
      for (char *t = strtok_r(dup, ",", &save); t && g_nothers < 64; t = strtok_r(NULL, ",", &save)) {
        unsigned pid = 0; unsigned long long b = 0; char kind[8] = "c"; int sm = 0;
        int got = sscanf(t, "%u:%llu:%7[a-z]:%d", &pid, &b, kind, &sm);
        if (got >= 2) {
          other_t *o = &g_others[g_nothers++];
          o->pid = pid; o->bytes = b; o->sm = sm;
          o->compute = strchr(kind, 'c') != NULL;
          o->graphics = strchr(kind, 'g') != NULL;
        }
      }
      free(dup);
    }
    if ((s = getenv("STUB_UTIL"))) {
      if (!strncmp(s, "fixed:", 6)) { g_util_mode = 0; g_util_fixed = atoi(s + 6); }
      else if (!strncmp(s, "closed:", 7)) { g_util_mode = 1; g_util_k = atof(s + 7); }
    }
    g_inited = 1;
  }
  pthread_mutex_unlock(&g_mu);
}

/* harness control (dlsym'd by the test programs) */
EXPORT void stub_ctl_set_util(int pct) { g_util_mode = 0; g_util_fixed = pct; }
EXPORT unsigned long long stub_ctl_launches(void) { return g_launches; }
EXPORT unsigned long long stub_ctl_device_bytes(int dev) { return g_dev_bytes[dev & 15]; }

static void track(void *p, size_t n, int kind, int dev) {
  pthread_mutex_lock(&g_mu);
  if (g_nallocs == g_callocs) {
    g_callocs = g_callocs ? g_callocs * 2 : 256;
    g_allocs = (alloc_t *)realloc(g_allocs, g_callocs * sizeof(alloc_t));
  }
  g_allocs[g_nallocs++] = (alloc_t){p, n, kind, dev};
  if (kind == 0 || kind == 3) g_dev_bytes[dev] += n;
  pthread_mutex_unlock(&g_mu);
}

static int untrack(void *p, alloc_t *out) {
  int found = 0;
  pthread_mutex_lock(&g_mu);
  for (size_t i = g_nallocs; i-- > 0;)
    if (g_allocs[i].p == p) {
      *out = g_allocs[i];
      g_allocs[i] = g_allocs[--g_nallocs];
      if (out->kind == 0 || out->kind == 3) g_dev_bytes[out->dev] -= out->n;
      found = 1;
      break;
    }
  pthread_mutex_unlock(&g_mu);
  return found;
}

static void *big_alloc(size_t n) {
  size_t len = (n + 4095) & ~(size_t)4095;
  if (!len) len = 4096;
  void *p = mmap(NULL, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  return p == MAP_FAILED ? NULL : p;
}
static void big_free(void *p, size_t n) {
  size_t len = (n + 4095) & ~(size_t)4095;
  if (!len) len = 4096;
  munmap(p, len);
}

/* ------------------------------------------------------------------ CUDA: basics */
EXPORT CUresult cuInit(unsigned f) { (void)f; stub_init(); return 0; }
EXPORT CUresult cuDriverGetVersion(int *v) { stub_init(); *v = 12090; return 0; }
EXPORT CUresult cuDeviceGetCount(int *n) { stub_init(); *n = g_gpu_count; return 0; }
EXPORT CUresult cuDeviceGet(CUdevice *d, int ord) { stub_init(); if (ord < 0 || ord >= g_gpu_count) return 101; *d = ord; return 0; }
EXPORT CUresult cuDeviceGetAttribute(int *pi, int attr, CUdevice d) {
  (void)d;
  switch (attr) {
  case 16: *pi = 148; break;     /* MULTIPROCESSOR_COUNT */
  case 39: *pi = 2048; break;    /* MAX_THREADS_PER_MULTIPROCESSOR */
  case 122: *pi = 1; break;      /* CAN_USE_64_BIT_STREAM_MEM_OPS */
  case 75: *pi = 10; break;      /* CC major */
  case 76: *pi = 0; break;
  default: *pi = 0;
  }
  return 0;
}
static void uuid_bytes(int dev, unsigned char *b) { memset(b, 0x11 * (dev + 1), 16); }
EXPORT CUresult cuDeviceGetUuid(void *uuid, CUdevice d) { uuid_bytes(d, (unsigned char *)uuid); return 0; }
EXPORT CUresult cuDeviceGetUuid_v2(void *uuid, CUdevice d) { uuid_bytes(d, (unsigned char *)uuid); return 0; }
EXPORT CUresult cuDeviceGetName(char *name, int len, CUdevice d) { snprintf(name, len, "STUB B200 #%d", d); return 0; }
EXPORT CUresult cuDeviceTotalMem_v2(size_t *b, CUdevice d) { (void)d; stub_init(); *b = g_total_mem; return 0; }
EXPORT CUresult cuGetErrorString(CUresult r, const char **s) {
  *s = r == 0 ? "no error" : r == 2 ? "out of memory" : r == 201 ? "invalid device context" : "stub error";
  return 0;
}
EXPORT CUresult cuGetErrorName(CUresult r, const char **s) { return cuGetErrorString(r, s); }

/* contexts: one implicit primary context per device, made current by the harness */
EXPORT CUresult cuDevicePrimaryCtxRetain(void **ctx, CUdevice d) { stub_init(); *ctx = (void *)(uintptr_t)(0x1000 + d); g_any_ctx = 1; g_ctx_on[d & 15] = 1; return 0; }
EXPORT CUresult cuDevicePrimaryCtxRelease_v2(CUdevice d) { (void)d; return 0; }
EXPORT CUresult cuDevicePrimaryCtxGetState(CUdevice d, unsigned *flags, int *active) { if (flags) *flags = 0; *active = g_ctx_on[d & 15]; return 0; }
/* context death: everything that lived in it is unmapped, so a library that still touches its old
 * device or pinned memory afterwards faults instead of silently reading stale bytes */
EXPORT CUresult cuDevicePrimaryCtxReset_v2(CUdevice d) {
  /* only this device's context dies: what lives in another GPU's context stays (a two-GPU sweep had the library's
   * runtime on the other device faulting on memory the stub had wrongly taken away) */
  pthread_mutex_lock(&g_mu);
  size_t keep = 0;
  for (size_t i = 0; i < g_nallocs; i++) {
    if (g_allocs[i].dev == d) big_free(g_allocs[i].p, g_allocs[i].n);
    else g_allocs[keep++] = g_allocs[i];
  }
  g_nallocs = keep;
  g_dev_bytes[d & 15] = 0;
  g_ctx_on[d & 15] = 0;
  g_any_ctx = 0;
  for (int i = 0; i < 16; i++) g_any_ctx |= g_ctx_on[i];
  pthread_mutex_unlock(&g_mu);
  if (t_cur_dev == d) t_has_ctx = 0;
  return 0;
}
EXPORT CUresult cuCtxCreate_v2(void **ctx, unsigned f, CUdevice d) { (void)f; stub_init(); *ctx = (void *)(uintptr_t)(0x1000 + d); t_cur_dev = d; t_has_ctx = 1; g_any_ctx = 1; g_ctx_on[d & 15] = 1; return 0; }
EXPORT CUresult cuCtxDestroy_v2(void *ctx) { (void)ctx; return 0; }
EXPORT CUresult cuCtxSetCurrent(void *ctx) { if (!ctx) { t_has_ctx = 0; return 0; } t_cur_dev = (int)((uintptr_t)ctx - 0x1000); t_has_ctx = 1; return 0; }
EXPORT CUresult cuCtxGetCurrent(void **ctx) { *ctx = t_has_ctx ? (void *)(uintptr_t)(0x1000 + t_cur_dev) : NULL; return 0; }
static __thread void *t_stack[8];
static __thread int t_sp;
EXPORT CUresult cuCtxPushCurrent_v2(void *ctx) { if (t_sp < 8) t_stack[t_sp++] = t_has_ctx ? (void *)(uintptr_t)(0x1000 + t_cur_dev) : NULL; return cuCtxSetCurrent(ctx); }
EXPORT CUresult cuCtxPopCurrent_v2(void **ctx) { if (ctx) cuCtxGetCurrent(ctx); void *prev = t_sp ? t_stack[--t_sp] : NULL; return cuCtxSetCurrent(prev); }
EXPORT CUresult cuCtxGetDevice(CUdevice *d) { if (!t_has_ctx) return 201; *d = t_cur_dev; return 0; }
EXPORT CUresult cuCtxSynchronize(void) { return t_has_ctx ? 0 : 201; }
EXPORT CUresult cuCtxGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return 0; }

/* ------------------------------------------------------------------ CUDA: memory */
static CUresult dev_alloc(CUdeviceptr *dptr, size_t n, int kind) {
  stub_init();
  if (!t_has_ctx) return 201;
  if (n == 0) return 1;
  if (kind == 0 && g_dev_bytes[t_cur_dev] + g_ctx_bytes + n > g_phys_mem) return 2;
  void *p = big_alloc(n);
  if (!p) return 2;
  track(p, n, kind, t_cur_dev);
  *dptr = (CUdeviceptr)(uintptr_t)p;
  return 0;
}
EXPORT CUresult cuMemAlloc_v2(CUdeviceptr *d, size_t n) { return dev_alloc(d, n, 0); }
EXPORT CUresult cuMemAllocManaged(CUdeviceptr *d, size_t n, unsigned f) { (void)f; return dev_alloc(d, n, 1); }
EXPORT CUresult cuMemAllocPitch_v2(CUdeviceptr *d, size_t *pitch, size_t w, size_t h, unsigned e) {
  (void)e;
  size_t p = (w + 511) & ~(size_t)511;
  CUresult r = dev_alloc(d, p * h, 0);
  if (r == 0) *pitch = p;
  return r;
}
EXPORT CUresult cuMemAllocAsync(CUdeviceptr *d, size_t n, void *s) { (void)s; return dev_alloc(d, n, 0); }
EXPORT CUresult cuMemAllocAsync_ptsz(CUdeviceptr *d, size_t n, void *s) { (void)s; return dev_alloc(d, n, 0); }
EXPORT CUresult cuMemAllocFromPoolAsync(CUdeviceptr *d, size_t n, void *pool, void *s) { (void)pool; (void)s; return dev_alloc(d, n, 0); }
EXPORT CUresult cuMemAllocFromPoolAsync_ptsz(CUdeviceptr *d, size_t n, void *pool, void *s) { (void)pool; (void)s; return dev_alloc(d, n, 0); }
static unsigned long long g_freed_dirty, g_freed_clean;
EXPORT unsigned long long stub_ctl_freed(int clean) { return clean ? g_freed_clean : g_freed_dirty; }
static CUresult dev_free(CUdeviceptr d) {
  if (!t_has_ctx) return 201;
  alloc_t a;
  if (!untrack((void *)(uintptr_t)d, &a)) return 1;
  if (a.kind == 0 && a.n >= 8 && getenv("STUB_CHECK_SCRUB")) {
    /* the harness dirtied the first and last bytes (scenario `dirty`); were they zeroed? */
    const unsigned char *p = (const unsigned char *)a.p;
    if (p[0] == 0 && p[a.n - 1] == 0) g_freed_clean++; else g_freed_dirty++;
    fprintf(stderr, "stub: freed %s buffer of %zu bytes\n", (p[0] == 0 && p[a.n - 1] == 0) ? "clean" : "dirty", a.n);
  }
  big_free(a.p, a.n);
  return 0;
}
EXPORT CUresult cuMemFree_v2(CUdeviceptr d) { return dev_free(d); }
EXPORT CUresult cuMemFreeAsync(CUdeviceptr d, void *s) { (void)s; return dev_free(d); }
EXPORT CUresult cuMemFreeAsync_ptsz(CUdeviceptr d, void *s) { (void)s; return dev_free(d); }
EXPORT CUresult cuMemGetInfo_v2(size_t *fr, size_t *tot) {
  stub_init();
  if (!t_has_ctx) return 201;
  uint64_t used = g_dev_bytes[t_cur_dev] + g_ctx_bytes;
  for (int i = 0; i < g_nothers; i++) used += g_others[i].bytes;
  *tot = g_total_mem;
  *fr = used >= g_total_mem ? 0 : g_total_mem - used;
  return 0;
}
/* ---- virtual memory management: handles are memfds, mappings are MAP_SHARED|MAP_FIXED views of them,
 * so re-pointing a virtual address at another handle behaves like the real thing (contents travel
 * with the handle, not with the address) */
typedef struct { int fd; size_t size; int host; } vmm_handle_t;
EXPORT CUresult cuMemGetAllocationGranularity(size_t *g, const void *prop, int opt) { (void)prop; (void)opt; *g = (size_t)2 << 20; return 0; }
EXPORT CUresult cuMemCreate(unsigned long long *h, size_t n, const void *prop, unsigned long long f) {
  (void)f;
  stub_init();
  const int *pi = (const int *)prop; /* CUmemAllocationProp: {type, requestedHandleTypes, location{type, id}, ...} */
  int host = pi && pi[2] != 1;
  int dev = t_has_ctx ? t_cur_dev : (pi && pi[2] == 1 ? pi[3] : 0);
  if (!host && g_dev_bytes[dev] + g_ctx_bytes + n > g_phys_mem) return 2;
  vmm_handle_t *v = (vmm_handle_t *)calloc(1, sizeof *v);
  v->fd = memfd_create("stub-vmm", 0);
  v->size = n;
  v->host = host;
  if (v->fd < 0 || ftruncate(v->fd, (off_t)n) != 0) { if (v->fd >= 0) close(v->fd); free(v); return 2; }
  track(v, n, host ? 4 : 3, dev);
  *h = (unsigned long long)(uintptr_t)v;
  return 0;
}
EXPORT CUresult cuMemRelease(unsigned long long h) {
  alloc_t a;
  if (!untrack((void *)(uintptr_t)h, &a)) return 1;
  vmm_handle_t *v = (vmm_handle_t *)a.p;
  close(v->fd);
  free(v);
  return 0;
}
EXPORT CUresult cuMemAddressReserve(CUdeviceptr *ptr, size_t size, size_t align, CUdeviceptr addr, unsigned long long flags) {
  (void)align; (void)addr; (void)flags;
  void *p = mmap(NULL, size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) return 2;
  *ptr = (CUdeviceptr)(uintptr_t)p;
  return 0;
}
EXPORT CUresult cuMemAddressFree(CUdeviceptr ptr, size_t size) { return munmap((void *)(uintptr_t)ptr, size) == 0 ? 0 : 1; }
EXPORT CUresult cuMemMap(CUdeviceptr ptr, size_t size, size_t offset, unsigned long long h, unsigned long long flags) {
  (void)flags;
  vmm_handle_t *v = (vmm_handle_t *)(uintptr_t)h;
  if (!v || offset + size > v->size) return 1;
  void *p = mmap((void *)(uintptr_t)ptr, size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, v->fd, (off_t)offset);
  return p == MAP_FAILED ? 1 : 0;
}
EXPORT CUresult cuMemUnmap(CUdeviceptr ptr, size_t size) {
  void *p = mmap((void *)(uintptr_t)ptr, size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED | MAP_NORESERVE, -1, 0);
  return p == MAP_FAILED ? 1 : 0;
}
EXPORT CUresult cuMemSetAccess(CUdeviceptr ptr, size_t size, const void *desc, size_t n) { (void)ptr; (void)size; (void)desc; (void)n; return 0; }
EXPORT unsigned long long stub_ctl_host_vmm_bytes(void) {
  unsigned long long n = 0;
  pthread_mutex_lock(&g_mu);
  for (size_t i = 0; i < g_nallocs; i++) if (g_allocs[i].kind == 4) n += g_allocs[i].n;
  pthread_mutex_unlock(&g_mu);
  return n;
}
/* events: host clock (the fake GPU executes synchronously) */
EXPORT CUresult cuEventCreate(void **e, unsigned f) { (void)f; *e = calloc(1, sizeof(struct timespec)); return *e ? 0 : 2; }
EXPORT CUresult cuEventRecord(void *e, void *s) { (void)s; clock_gettime(CLOCK_MONOTONIC, (struct timespec *)e); return 0; }
EXPORT CUresult cuEventSynchronize(void *e) { (void)e; return 0; }
EXPORT CUresult cuEventElapsedTime(float *ms, void *a, void *b) {
  struct timespec *x = (struct timespec *)a, *y = (struct timespec *)b;
  *ms = (float)((y->tv_sec - x->tv_sec) * 1e3 + (y->tv_nsec - x->tv_nsec) * 1e-6);
  return 0;
}
EXPORT CUresult cuEventDestroy_v2(void *e) { free(e); return 0; }
typedef struct { size_t W, H; int fmt; unsigned ch; } arr2_t;
typedef struct { size_t W, H, D; int fmt; unsigned ch, flags; } arr3_t;
static size_t fmt_bytes(int f) { return (f == 1 || f == 8) ? 1 : (f == 2 || f == 9 || f == 0x10) ? 2 : 4; }
EXPORT CUresult cuArrayCreate_v2(void **h, const arr2_t *d) {
  CUdeviceptr p;
  CUresult r = dev_alloc(&p, fmt_bytes(d->fmt) * d->ch * d->W * (d->H ? d->H : 1), 0);
  if (r == 0) *h = (void *)(uintptr_t)p;
  return r;
}
EXPORT CUresult cuArray3DCreate_v2(void **h, const arr3_t *d) {
  CUdeviceptr p;
  CUresult r = dev_alloc(&p, fmt_bytes(d->fmt) * d->ch * d->W * (d->H ? d->H : 1) * (d->D ? d->D : 1), 0);
  if (r == 0) *h = (void *)(uintptr_t)p;
  return r;
}
EXPORT CUresult cuMipmappedArrayCreate(void **h, const arr3_t *d, unsigned levels) { (void)levels; return cuArray3DCreate_v2(h, d); }
EXPORT CUresult cuArrayDestroy(void *h) { return dev_free((CUdeviceptr)(uintptr_t)h); }
EXPORT CUresult cuMipmappedArrayDestroy(void *h) { return dev_free((CUdeviceptr)(uintptr_t)h); }
EXPORT CUresult cuMemHostAlloc(void **pp, size_t n, unsigned f) { (void)f; *pp = big_alloc(n); if (!*pp) return 2; track(*pp, n, 2, t_has_ctx ? t_cur_dev : 0); return 0; }
EXPORT CUresult cuMemFreeHost(void *p) { alloc_t a; if (!untrack(p, &a)) return 1; big_free(a.p, a.n); return 0; }
EXPORT CUresult cuMemHostGetDevicePointer_v2(CUdeviceptr *d, void *p, unsigned f) { (void)f; *d = (CUdeviceptr)(uintptr_t)p; return 0; }
EXPORT CUresult cuMemGetAddressRange_v2(CUdeviceptr *base, size_t *size, CUdeviceptr d) {
  CUresult r = 1;
  pthread_mutex_lock(&g_mu);
  for (size_t i = 0; i < g_nallocs; i++)
    if ((uintptr_t)g_allocs[i].p <= (uintptr_t)d && (uintptr_t)d < (uintptr_t)g_allocs[i].p + g_allocs[i].n) {
      if (base) *base = (CUdeviceptr)(uintptr_t)g_allocs[i].p;
      if (size) *size = g_allocs[i].n;
      r = 0;
      break;
    }
  pthread_mutex_unlock(&g_mu);
  return r;
}
EXPORT CUresult cuMemsetD8_v2(CUdeviceptr d, unsigned char v, size_t n) { memset((void *)(uintptr_t)d, v, n); return 0; }
EXPORT CUresult cuMemcpyDtoH_v2(void *dst, CUdeviceptr src, size_t n) { memcpy(dst, (void *)(uintptr_t)src, n); return 0; }
EXPORT CUresult cuMemcpyDtoHAsync_v2(void *dst, CUdeviceptr src, size_t n, void *s) { (void)s; memcpy(dst, (void *)(uintptr_t)src, n); return 0; }
EXPORT CUresult cuMemcpyHtoD_v2(CUdeviceptr dst, const void *src, size_t n) { memcpy((void *)(uintptr_t)dst, src, n); return 0; }
EXPORT CUresult cuMemcpyDtoD_v2(CUdeviceptr dst, CUdeviceptr src, size_t n) { memmove((void *)(uintptr_t)dst, (void *)(uintptr_t)src, n); return 0; }
EXPORT CUresult cuMemcpy(CUdeviceptr dst, CUdeviceptr src, size_t n) { memmove((void *)(uintptr_t)dst, (void *)(uintptr_t)src, n); return 0; }

/* STUB_CTX_LOCK=1 models the real driver's context lock: a call that blocks on the stream (here
 * the wait itself, on a real GPU e.g. a pageable memcpy behind a parked kernel) keeps every other
 * thread out of the driver - in particular out of cuLaunchKernel. */
/* FIFO (ticket) lock: a plain pthread mutex lets the releasing thread re-acquire it at once and
 * would starve the library's tick thread for ever, which is a property of the model, not of
 * the thing modelled */
static pthread_mutex_t g_ctx_m = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_ctx_c = PTHREAD_COND_INITIALIZER;
static unsigned long g_ctx_next, g_ctx_serving;
static void ctx_lock(void) {
  pthread_mutex_lock(&g_ctx_m);
  unsigned long my = g_ctx_next++;
  while (my != g_ctx_serving) pthread_cond_wait(&g_ctx_c, &g_ctx_m);
  pthread_mutex_unlock(&g_ctx_m);
}
static void ctx_unlock(void) {
  pthread_mutex_lock(&g_ctx_m);
  g_ctx_serving++;
  pthread_cond_broadcast(&g_ctx_c);
  pthread_mutex_unlock(&g_ctx_m);
}
static int g_ctx_lock = -1;
static CUresult wait64_unlocked(CUdeviceptr addr, unsigned long long value);
static __thread void *t_wait_stream;
static int ctx_lock_on(void) {
  if (g_ctx_lock < 0) { const char *e = getenv("STUB_CTX_LOCK"); g_ctx_lock = (e && atoi(e)) ? 1 : 0; }
  return g_ctx_lock;
}
/* ------------------------------------------------------------------ CUDA: streams */
EXPORT CUresult cuStreamCreate(void **s, unsigned f) { (void)f; static uintptr_t next = 0x5000; *s = (void *)(next += 0x10); return 0; }
EXPORT CUresult cuStreamCreateWithPriority(void **s, unsigned f, int p) { (void)p; return cuStreamCreate(s, f); }
EXPORT CUresult cuStreamDestroy_v2(void *s) { (void)s; return 0; }
EXPORT CUresult cuStreamSynchronize(void *s) { (void)s; return 0; }
EXPORT CUresult cuStreamQuery(void *s) {
  (void)s;
  if (ctx_lock_on()) { ctx_lock(); ctx_unlock(); }
  for (int i = 0; i < 64; i++)
    if (g_parked_streams[i] == (s ? s : (void *)1)) return 600; /* NOT_READY: a stream-wait is pending on this stream */
  return 0;
}
EXPORT CUresult cuStreamIsCapturing(void *s, int *st) { (void)s; *st = 0; return 0; }
EXPORT CUresult cuThreadExchangeStreamCaptureMode(int *mode) { static __thread int cur = 0; int old = cur; cur = *mode; *mode = old; return 0; }
static CUresult wait64_unlocked(CUdeviceptr addr, unsigned long long value);
static CUresult wait64(CUdeviceptr addr, unsigned long long value) {
  if (!ctx_lock_on()) return wait64_unlocked(addr, value);
  ctx_lock();
  CUresult r = wait64_unlocked(addr, value);
  ctx_unlock();
  return r;
}
static CUresult wait64_unlocked(CUdeviceptr addr, unsigned long long value) {
  /* the fake GPU executes stream work synchronously, so a stream wait blocks the caller */
  volatile long long *p = (volatile long long *)(uintptr_t)addr;
  struct timespec nap = {0, 200000};
  if ((long long)(*p - (long long)value) >= 0) return 0;
  __sync_fetch_and_add(&g_parked, 1);
  void *tag = t_wait_stream ? t_wait_stream : (void *)1; /* 1 = the legacy stream */
  int mine = -1;
  for (int i = 0; i < 64 && mine < 0; i++)
    if (__sync_bool_compare_and_swap(&g_parked_streams[i], NULL, tag)) mine = i;
  for (int i = 0; i < 100000; i++) { /* 20 s cap */
    if ((long long)(*p - (long long)value) >= 0) break;
    nanosleep(&nap, NULL);
  }
  if (mine >= 0) g_parked_streams[mine] = NULL;
  __sync_fetch_and_sub(&g_parked, 1);
  return 0;
}
EXPORT CUresult cuStreamWaitValue64_v2(void *s, CUdeviceptr a, unsigned long long v, unsigned f) { (void)f; t_wait_stream = s; return wait64(a, v); }
EXPORT CUresult cuStreamWaitValue64_v2_ptsz(void *s, CUdeviceptr a, unsigned long long v, unsigned f) { (void)f; t_wait_stream = s; return wait64(a, v); }
EXPORT CUresult cuStreamWriteValue64_v2(void *s, CUdeviceptr a, unsigned long long v, unsigned f) { (void)s; (void)f; *(volatile unsigned long long *)(uintptr_t)a = v; return 0; }
EXPORT CUresult cuStreamWriteValue64_v2_ptsz(void *s, CUdeviceptr a, unsigned long long v, unsigned f) { (void)s; (void)f; *(volatile unsigned long long *)(uintptr_t)a = v; return 0; }

/* ------------------------------------------------------------------ CUDA: modules and the fake GPU's kernels */
typedef struct { char name[64]; } fn_t;
/* STUB_FAIL_MODULE=N: the first N module loads fail like a full GPU (the library's bring-up must free what it built,
 * back off and try again) */
EXPORT CUresult cuModuleLoadData(void **m, const void *img) {
  static int left = -1;
  (void)img;
  if (left < 0) { const char *s = getenv("STUB_FAIL_MODULE"); left = s ? atoi(s) : 0; }
  if (left > 0) { left--; return 2; }
  *m = (void *)0x7000;
  return 0;
}
EXPORT CUresult cuModuleUnload(void *m) { (void)m; return 0; }
EXPORT CUresult cuModuleGetFunction(void **f, void *m, const char *name) {
  (void)m;
  fn_t *fn = (fn_t *)calloc(1, sizeof(fn_t));
  snprintf(fn->name, sizeof fn->name, "%s", name);
  *f = fn;
  return 0;
}
EXPORT CUresult cuFuncSetAttribute(void *f, int a, int v) { (void)f; (void)a; (void)v; return 0; }
EXPORT CUresult cuFuncSetBlockShape(void *f, int x, int y, int z) { (void)f; (void)x; (void)y; (void)z; return 0; }

static void fake_quota(const vgpu_quota_req_t *q, vgpu_quota_res_t *r) {
  uint8_t cp[VGPU_MAX_PIDS], cl[VGPU_MAX_PIDS], gp[VGPU_MAX_PIDS], gl[VGPU_MAX_PIDS];
  for (uint32_t i = 0; i < q->n_compute; i++) { cp[i] = q->cflags[i] & VGPU_FLAG_PRIMARY; cl[i] = (q->cflags[i] & VGPU_FLAG_LOCAL) != 0; }
  for (uint32_t i = 0; i < q->n_graphics; i++) { gp[i] = q->gflags[i] & VGPU_FLAG_PRIMARY; gl[i] = (q->gflags[i] & VGPU_FLAG_LOCAL) != 0; }
  uint64_t used = orc_used_memory((int)q->mode, q->compute, q->n_compute, cp, cl, q->graphics, q->n_graphics, gp, gl);
  /* own-footprint compensation, same rule as the kernel: skip it when our own record is
   * visible in the lists but was not counted as a container member */
  int seen = 0;
  vgpu_proc_t *c2 = (vgpu_proc_t *)malloc(sizeof(vgpu_proc_t) * VGPU_MAX_PIDS * 2), *g2 = c2 + VGPU_MAX_PIDS;
  memcpy(c2, q->compute, sizeof(vgpu_proc_t) * q->n_compute);
  memcpy(g2, q->graphics, sizeof(vgpu_proc_t) * q->n_graphics);
  for (uint32_t i = 0; i < q->n_compute; i++) if (c2[i].pid == q->self_pid) { seen = 1; c2[i].used_bytes = c2[i].used_bytes ? 0 : 1; }
  for (uint32_t i = 0; i < q->n_graphics; i++) if (g2[i].pid == q->self_pid) { seen = 1; g2[i].used_bytes = g2[i].used_bytes ? 0 : 1; }
  uint64_t without = orc_used_memory((int)q->mode, c2, q->n_compute, cp, cl, g2, q->n_graphics, gp, gl);
  free(c2);
  int counted = without != used;
  uint64_t self = (seen && !counted) ? 0 : q->self_bytes;
  used = used >= self ? used - self : 0;
  vgpu_vmem_dev_t *led = (vgpu_vmem_dev_t *)calloc(1, sizeof *led);
  memcpy(led->processes, q->vmem, (size_t)q->n_vmem * sizeof(vgpu_vmem_rec_t));
  led->processes_size = q->n_vmem;
  uint64_t vmem = orc_ledger_sum(led);
  free(led);
  vgpu_cfg_dev_t c;
  memset(&c, 0, sizeof c);
  c.total_memory = q->total_memory;
  c.real_memory = q->real_memory;
  c.memory_oversold = (int)q->memory_oversold;
  c.memory_limit = 1;
  r->used = used;
  r->vmem = vmem;
  r->total = q->total_memory;
  r->out_used = r->out_free = 0;
  r->path = VGPU_PATH_GPU;
  if (q->kind == VGPU_Q_ALLOC) {
    r->path = (uint32_t)orc_memory_path(&c, used, vmem, q->request, (int)q->allow_uva);
  } else if (q->kind == VGPU_Q_NVML_INFO) {
    orc_nvml_meminfo(&c, used, vmem, &r->total, &r->out_used, &r->out_free);
  } else {
    orc_cu_meminfo(&c, used, vmem, (int)q->real_ok, q->real_total, &r->out_free, &r->total);
    r->out_used = r->total - r->out_free;
  }
  __sync_synchronize();
  r->seq_done = q->seq;
}

typedef struct { const vgpu_quota_req_t *q; vgpu_quota_res_t *r; uint32_t seq; } armed_arg_t;
static void *fake_armed_quota(void *argp) {
  armed_arg_t a = *(armed_arg_t *)argp;
  free(argp);
  struct timespec nap = {0, 20000};
  int ok = 0;
  for (int i = 0; i < 1000 && !ok; i++) { /* 20 ms, like the kernel */
    ok = *(volatile uint32_t *)&a.q->seq == a.seq;
    if (!ok) nanosleep(&nap, NULL);
  }
  __sync_synchronize();
  if (ok) {
    fake_quota(a.q, a.r);
  } else {
    a.r->path = VGPU_PATH_RETRY;
    __sync_synchronize();
    a.r->seq_done = a.seq;
  }
  return NULL;
}

static void fake_ctl_step(vgpu_lim_dev_t *D, vgpu_lim_host_t *H, int user, int sys, int valid, int nproc) {
  orc_gpu_t g = {D->sm_num, D->max_thread_per_sm, D->total_cores};
  vgpu_cfg_dev_t c;
  memset(&c, 0, sizeof c);
  c.hard_core = D->hard_core; c.soft_core = D->soft_core; c.core_limit = D->core_limit; c.hard_limit = D->hard_limit;
  orc_watcher_t w = {D->share, D->sys_free, D->avg_sys_free, D->ctr_i, D->pre_sys_process_num, D->up_limit, 0};
  if (valid) D->valid = 1;
  orc_util_t u = {user, sys, D->valid, nproc};
  if (H->release_pending) {
    long long fl = H->release_floor;
    H->release_pending = 0;
    if (fl - D->granted > 0) D->granted = fl;
  }
  long long consumed = H->consumed;
  int64_t bucket = D->granted - consumed;
  if (H->ext_limits_seq != D->limits_seen) { /* node-level rebalance, as in ctl_step */
    D->limits_seen = H->ext_limits_seq;
    int soft = H->ext_soft_core, up = H->ext_up_limit;
    if (D->limits_seen != 0 && soft > D->hard_core) {
      D->soft_core = soft; D->hard_limit = 0;
      c.soft_core = soft; c.hard_limit = 0;
      D->ext_up = up < D->hard_core ? D->hard_core : (up > soft ? soft : up);
    } else {
      D->ext_up = 0;
    }
  }
  if (D->ext_up > 0 && !D->hard_limit && D->core_limit && D->valid) {
    w.sys_free = 100 - sys;
    w.up_limit = D->ext_up;
    w.share = orc_delta(&g, w.up_limit, user, w.share);
    bucket = orc_change_token(&g, bucket, w.share);
  } else {
    orc_watcher_step(&g, &c, &w, &u, &bucket);
  }
  D->share = w.share; D->sys_free = w.sys_free; D->avg_sys_free = w.avg_sys_free; D->ctr_i = w.i;
  D->pre_sys_process_num = w.pre_sys_process_num; D->up_limit = w.up_limit;
  if (D->core_limit && D->valid) D->granted = bucket + consumed;
  D->bucket_last = bucket;
  D->last_user_current = user; D->last_sys_current = sys;
  D->steps++;
  H->granted_mirror = D->granted; H->bucket_mirror = bucket; H->share_mirror = D->share;
  H->up_limit_mirror = D->up_limit; H->user_current = user; H->sys_current = sys;
  H->steps = D->steps;
}

static void fake_period_end(vgpu_lim_dev_t *D, vgpu_lim_host_t *H) {
  int q = D->total_samples ? (int)(D->busy_samples * 100 / D->total_samples) : 0;
  D->last_queue_busy_pct = q;
  D->busy_samples = D->total_samples = 0;
  int user = H->ext_user_override >= 0 ? H->ext_user_override : q;
  int others = H->ext_sys_current > 0 ? H->ext_sys_current : 0;
  int np = H->ext_sys_process_num > 0 ? H->ext_sys_process_num : 1;
  fake_ctl_step(D, H, user, user + others, 1, np);
}

/* vgpu_governor_kernel on the fake GPU: the only kernel that is not executed synchronously -
 * it is resident, so it runs as a thread until its own retire protocol lets it go. */
typedef struct { vgpu_lim_dev_t *D; vgpu_lim_host_t *H; uint32_t interval_us, period_us, idle_us; } gov_arg_t;
static uint64_t mono_ns(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
static unsigned long long fake_signature(const vgpu_lim_host_t *H) {
  unsigned long long sgn = 0;
  for (uint32_t i = 0; i < VGPU_STREAM_SLOTS; i++) sgn += H->launched[i] * 3ull + H->done[i];
  return sgn;
}
static void *fake_governor(void *argp) {
  gov_arg_t a = *(gov_arg_t *)argp;
  free(argp);
  vgpu_lim_dev_t *D = a.D;
  vgpu_lim_host_t *H = a.H;
  const uint64_t period_ns = (uint64_t)(a.period_us ? a.period_us : 1) * 1000ull, idle_ns = (uint64_t)a.idle_us * 1000ull;
  uint64_t now = mono_ns();
  uint64_t left = D->gov_left_ns, lc = D->last_ctl_ns;
  if (lc == 0 || now < lc || left == 0 || now < left || left < lc) {
    D->last_ctl_ns = now;
  } else {
    int was_busy = D->gov_left_busy != 0;
    uint64_t t = left;
    unsigned steps = 0;
    while (now - lc >= period_ns && steps < 128) {
      uint64_t end = lc + period_ns;
      D->total_samples += end - t;
      if (was_busy) D->busy_samples += end - t;
      fake_period_end(D, H);
      t = lc = end;
      steps++;
    }
    if (now - lc >= period_ns) { lc = now - (now - lc) % period_ns; t = lc; }
    D->total_samples += now - t;
    if (was_busy) D->busy_samples += now - t;
    D->last_ctl_ns = lc;
  }
  D->gov_left_ns = 0;
  D->gov_left_busy = 0;
  H->gov_left_busy = 0;
  uint64_t prev = now, last_change = now;
  unsigned long long sig = fake_signature(H), busy = 0, total = 0;
  unsigned nap_us = a.interval_us < 200 ? 200 : a.interval_us; /* a CPU thread: do not spin at 50 us */
  for (;;) {
    now = mono_ns();
    uint64_t dt = now - prev;
    prev = now;
    int util = current_util();
    int parked = g_parked > 0, outstanding = 0;
    for (uint32_t i = 0; i < VGPU_STREAM_SLOTS; i++)
      if (H->launched[i] > H->done[i]) outstanding = 1;
    total += dt;
    busy += dt * (unsigned long long)util / 100ull;
    unsigned long long sn = fake_signature(H);
    if (sn != sig) { sig = sn; last_change = now; }
    if (now - D->last_ctl_ns >= period_ns) {
      D->last_ctl_ns = now;
      D->busy_samples += busy;
      D->total_samples += total;
      busy = total = 0;
      fake_period_end(D, H);
    }
    uint32_t quit = H->quit;
    int want_exit = quit ? !parked : (!outstanding && !parked && now - last_change > idle_ns);
    if (want_exit) {
      H->ctl_state = 2;
      __sync_synchronize();
      unsigned long long chk = fake_signature(H);
      if (chk != sig && !quit) {
        H->ctl_state = 1;
        sig = chk;
        last_change = now;
      } else {
        D->busy_samples += busy;
        D->total_samples += total;
        D->gov_left_ns = now;
        D->gov_left_busy = 0;
        H->gov_left_busy = 0;
        __sync_synchronize();
        H->ctl_state = 0;
        return NULL;
      }
    }
    struct timespec nap = {0, (long)nap_us * 1000L};
    nanosleep(&nap, NULL);
  }
}

static void run_fake_kernel(const char *name, void **p) {
  if (!strcmp(name, VGPU_K_CLEAR)) {
    unsigned long long n = *(unsigned long long *)p[1];
    if (n) memset((void *)(uintptr_t) * (CUdeviceptr *)p[0], 0, n);
  } else if (!strcmp(name, VGPU_K_SPILL) || !strcmp(name, "vgpu_copy_generic_kernel")) {
    unsigned long long n = *(unsigned long long *)p[2];
    if (n) memmove((void *)(uintptr_t) * (CUdeviceptr *)p[0], (void *)(uintptr_t) * (CUdeviceptr *)p[1], n);
  } else if (!strcmp(name, VGPU_K_QUOTA)) {
    const vgpu_quota_req_t *q = (const vgpu_quota_req_t *)(uintptr_t) * (CUdeviceptr *)p[0];
    vgpu_quota_res_t *r = (vgpu_quota_res_t *)(uintptr_t) * (CUdeviceptr *)p[1];
    uint32_t armed = *(uint32_t *)p[2];
    if (!armed) {
      fake_quota(q, r);
    } else {
      /* armed launch: the real kernel is asynchronous and waits on the device for the request to be
       * published under `armed`; the synchronous fake GPU needs a thread for that */
      armed_arg_t *a = (armed_arg_t *)malloc(sizeof *a);
      a->q = q; a->r = r; a->seq = armed;
      pthread_t t;
      pthread_attr_t at;
      pthread_attr_init(&at);
      pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
      if (pthread_create(&t, &at, fake_armed_quota, a) != 0) { r->path = VGPU_PATH_RETRY; __sync_synchronize(); r->seq_done = armed; free(a); }
      pthread_attr_destroy(&at);
    }
  } else if (!strcmp(name, VGPU_K_SLAB_INSERT)) {
    vgpu_slab_slot_t *slab = (vgpu_slab_slot_t *)(uintptr_t) * (CUdeviceptr *)p[0];
    unsigned long long dptr = *(unsigned long long *)p[1], bytes = *(unsigned long long *)p[2];
    vgpu_slab_res_t *res = (vgpu_slab_res_t *)(uintptr_t) * (CUdeviceptr *)p[3];
    uint32_t seq = *(uint32_t *)p[4], slot = 0xffffffffu;
    for (uint32_t i = 0; i < VGPU_SLAB_SLOTS; i++)
      if (slab[i].dptr <= 1) { slab[i].dptr = dptr; slab[i].bytes = bytes; slot = i; break; }
    res->bytes = bytes; res->slot = slot;
    __sync_synchronize();
    res->seq_done = seq;
  } else if (!strcmp(name, VGPU_K_SLAB_REMOVE)) {
    vgpu_slab_slot_t *slab = (vgpu_slab_slot_t *)(uintptr_t) * (CUdeviceptr *)p[0];
    unsigned long long dptr = *(unsigned long long *)p[1];
    vgpu_slab_res_t *res = (vgpu_slab_res_t *)(uintptr_t) * (CUdeviceptr *)p[2];
    uint32_t seq = *(uint32_t *)p[3], slot = 0xffffffffu;
    unsigned long long bytes = 0;
    for (uint32_t i = 0; i < VGPU_SLAB_SLOTS; i++)
      if (slab[i].dptr == dptr && dptr > 1) { bytes = slab[i].bytes; slab[i].dptr = 1; slab[i].bytes = 0; slot = i; break; }
    res->bytes = bytes; res->slot = slot;
    __sync_synchronize();
    res->seq_done = seq;
  } else if (!strcmp(name, VGPU_K_VSLAB)) {
    vgpu_vslab_slot_t *tab = (vgpu_vslab_slot_t *)(uintptr_t) * (CUdeviceptr *)p[0];
    const vgpu_vslab_req_t *rq = (const vgpu_vslab_req_t *)p[1];
    vgpu_vslab_res_t *res = (vgpu_vslab_res_t *)(uintptr_t) * (CUdeviceptr *)p[2];
    uint32_t seq = *(uint32_t *)p[3], slot = 0xffffffffu;
    unsigned long long best = ~0ull;
    for (uint32_t i = 0; i < VGPU_VSLAB_SLOTS; i++) {
      unsigned long long k = ~0ull;
      if (rq->op == VGPU_VSLAB_PUT) { if (tab[i].dptr == rq->dptr) k = i; else if (tab[i].dptr == 0) k = (1ull << 32) | i; }
      else if (rq->op == VGPU_VSLAB_TAKE) { if (tab[i].dptr == rq->dptr && rq->dptr) k = i; }
      else if (tab[i].dptr && (rq->dptr ? tab[i].dptr == rq->dptr : (tab[i].size == rq->size && (tab[i].state & rq->mask) == rq->want)))
        k = ((unsigned long long)tab[i].age << 32) | i;
      if (k < best) best = k;
    }
    if (best != ~0ull) slot = (uint32_t)(best & 0xffffffffu);
    memset(res, 0, sizeof *res);
    res->slot = slot;
    if (slot != 0xffffffffu) {
      if (rq->op == VGPU_VSLAB_PUT) {
        tab[slot].dptr = rq->dptr; tab[slot].bytes = rq->bytes; tab[slot].size = rq->size; tab[slot].state = rq->state; tab[slot].age = rq->age;
      }
      res->dptr = tab[slot].dptr; res->bytes = tab[slot].bytes; res->size = tab[slot].size; res->state = tab[slot].state; res->age = tab[slot].age;
      if (rq->op == VGPU_VSLAB_TAKE) { tab[slot].dptr = 0; tab[slot].state = 0; }
      else if (rq->op == VGPU_VSLAB_SCAN && rq->set_mask) tab[slot].state = (res->state & ~rq->set_mask) | (rq->set_val & rq->set_mask);
    }
    __sync_synchronize();
    res->seq_done = seq;
  } else if (!strcmp(name, VGPU_K_CONTROLLER)) {
    vgpu_ctrl_in_t *in = (vgpu_ctrl_in_t *)p[2];
    fake_ctl_step((vgpu_lim_dev_t *)(uintptr_t) * (CUdeviceptr *)p[0], (vgpu_lim_host_t *)(uintptr_t) * (CUdeviceptr *)p[1],
                  in->user_current, in->sys_current, in->valid, in->sys_process_num);
  } else if (!strcmp(name, VGPU_K_REFILL)) {
    /* vgpu_refill_kernel: fold of the published samples (oracle restatement of cuda_hook.c:1044-1159)
     * into the persistent top_result, then one watcher step */
    vgpu_lim_dev_t *D = (vgpu_lim_dev_t *)(uintptr_t) * (CUdeviceptr *)p[0];
    vgpu_lim_host_t *H = (vgpu_lim_host_t *)(uintptr_t) * (CUdeviceptr *)p[1];
    const vgpu_util_req_t *U = (const vgpu_util_req_t *)(uintptr_t) * (CUdeviceptr *)p[2];
    if (U->status == VGPU_UTIL_SAMPLES && D->core_limit) {
      uint8_t prim[VGPU_MAX_PIDS], loc[VGPU_MAX_PIDS];
      uint32_t n = U->n_samples > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : U->n_samples;
      for (uint32_t i = 0; i < n; i++) { prim[i] = (U->flags[i] & VGPU_FLAG_PRIMARY) != 0; loc[i] = (U->flags[i] & VGPU_FLAG_LOCAL) != 0; }
      orc_util_t u = {0, 0, D->valid, 0};
      orc_fold_utilization((int)U->mode, U->samples, n, U->checktime_us, prim, loc, (int)U->have_container_pids, &u);
      D->top_user = u.user_current; D->top_sys = u.sys_current;
      if (u.valid) D->valid = 1;
    }
    if (U->status != VGPU_UTIL_NOTHING) {
      int np = U->sys_process_num;
      if (U->status == VGPU_UTIL_SAMPLES && (U->mode & VGPU_MODE_OPEN_KERNEL) == VGPU_MODE_OPEN_KERNEL && (int)U->n_samples > np) np = (int)U->n_samples;
      D->top_nproc = np;
    }
    D->top_seq = (int)U->seq;
    int user = H->ext_user_override >= 0 ? H->ext_user_override : D->top_user;
    fake_ctl_step(D, H, user, D->top_sys, H->ext_user_override >= 0, D->top_nproc);
  } else if (!strcmp(name, VGPU_K_SAMPLER)) {
    vgpu_lim_dev_t *D = (vgpu_lim_dev_t *)(uintptr_t) * (CUdeviceptr *)p[0];
    vgpu_lim_host_t *H = (vgpu_lim_host_t *)(uintptr_t) * (CUdeviceptr *)p[1];
    uint32_t period = *(uint32_t *)p[4];
    uint32_t skipped = *(uint32_t *)p[6];
    int util = current_util();
    D->busy_samples += (unsigned long long)util;
    D->total_samples += 100ull * (1ull + skipped); /* skipped ticks = idle windows */
    if (period != VGPU_SAMPLER_PROBE_ONLY) {
      unsigned long long tick = (unsigned long long)D->period_tick + skipped + 1ull;
      if (tick < period) {
        D->period_tick = (uint32_t)tick;
      } else {
        unsigned long long periods = period ? tick / period : 1ull;
        D->period_tick = period ? (uint32_t)(tick % period) : 0u;
        fake_period_end(D, H);
        if (periods > 129ull) periods = 129ull;
        for (unsigned long long k = 1; k < periods; k++) {
          D->total_samples = 1;
          fake_period_end(D, H);
        }
      }
    }
  } else if (!strcmp(name, VGPU_K_GOVERNOR)) {
    gov_arg_t *a = (gov_arg_t *)malloc(sizeof *a);
    a->D = (vgpu_lim_dev_t *)(uintptr_t) * (CUdeviceptr *)p[0];
    a->H = (vgpu_lim_host_t *)(uintptr_t) * (CUdeviceptr *)p[1];
    a->interval_us = *(uint32_t *)p[2]; a->period_us = *(uint32_t *)p[3]; a->idle_us = *(uint32_t *)p[4];
    pthread_t t;
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
    if (pthread_create(&t, &at, fake_governor, a) != 0) { a->H->ctl_state = 0; free(a); }
    pthread_attr_destroy(&at);
  } else if (!strcmp(name, VGPU_K_GATE)) {
    wait64_unlocked(*(CUdeviceptr *)p[0], (unsigned long long)*(long long *)p[1]);
  } else {
    note_launch(); /* a tenant kernel */
  }
}

static CUresult launch(void *f, void **params) {
  if (!t_has_ctx) return 201;
  fn_t *fn = (fn_t *)f;
  int locked = ctx_lock_on();
  if (locked) ctx_lock();
  if (fn && ((uintptr_t)fn > 0x10000) && !strncmp(fn->name, "vgpu_", 5)) run_fake_kernel(fn->name, params);
  else note_launch();
  if (locked) ctx_unlock();
  return 0;
}
EXPORT CUresult cuLaunchKernel(void *f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                               unsigned sm, void *s, void **params, void **extra) {
  (void)gx; (void)gy; (void)gz; (void)bx; (void)by; (void)bz; (void)sm; (void)s; (void)extra;
  return launch(f, params);
}
EXPORT CUresult cuLaunchKernel_ptsz(void *f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                    unsigned sm, void *s, void **params, void **extra) {
  /* not through the exported cuLaunchKernel: a preloaded library interposes that symbol, the real driver does not
   * call its own entry points through the PLT */
  (void)gx; (void)gy; (void)gz; (void)bx; (void)by; (void)bz; (void)sm; (void)s; (void)extra;
  return launch(f, params);
}
EXPORT CUresult cuLaunchKernelEx(const void *cfg, void *f, void **params, void **extra) { (void)cfg; (void)extra; return launch(f, params); }
EXPORT CUresult cuLaunchKernelEx_ptsz(const void *cfg, void *f, void **params, void **extra) { (void)cfg; (void)extra; return launch(f, params); }
EXPORT CUresult cuLaunchCooperativeKernel(void *f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz,
                                          unsigned sm, void *s, void **params) {
  (void)gx; (void)gy; (void)gz; (void)bx; (void)by; (void)bz; (void)sm; (void)s;
  return launch(f, params);
}
EXPORT CUresult cuLaunchCooperativeKernel_ptsz(void *f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                                               unsigned bz, unsigned sm, void *s, void **params) {
  (void)gx; (void)gy; (void)gz; (void)bx; (void)by; (void)bz; (void)sm; (void)s;
  return launch(f, params);
}
EXPORT CUresult cuLaunch(void *f) { return launch(f, NULL); }
EXPORT CUresult cuLaunchGrid(void *f, int w, int h) { (void)w; (void)h; return launch(f, NULL); }
EXPORT CUresult cuLaunchGridAsync(void *f, int w, int h, void *s) { (void)w; (void)h; (void)s; return launch(f, NULL); }

/* cuGetProcAddress: hand out this object's own symbols.  A static table, not dlsym(): under
 * LD_PRELOAD the process-wide dlsym is the interposer of the library under test. */
/* ------------------------------------------------------------------ CUDA: graphs (kernel nodes only) */
typedef struct { void *func; unsigned gx, gy, gz, bx, by, bz, smem; void **params; void **extra; } knode_params_t;
typedef struct fake_graph { int n; knode_params_t nodes[64]; } fake_graph_t;
typedef struct { fake_graph_t g; } fake_exec_t;
EXPORT CUresult cuGraphCreate(void **g, unsigned flags) { (void)flags; *g = calloc(1, sizeof(fake_graph_t)); return *g ? 0 : 2; }
EXPORT CUresult cuGraphDestroy(void *g) { free(g); return 0; }
EXPORT CUresult cuGraphAddKernelNode(void **node, void *g, const void **deps, size_t ndeps, const knode_params_t *p) {
  (void)deps; (void)ndeps;
  fake_graph_t *fg = (fake_graph_t *)g;
  if (!fg || fg->n >= 64) return 1;
  fg->nodes[fg->n] = *p;
  if (node) *node = &fg->nodes[fg->n];
  fg->n++;
  return 0;
}
EXPORT CUresult cuGraphGetNodes(void *g, void **nodes, size_t *n) {
  fake_graph_t *fg = (fake_graph_t *)g;
  if (!nodes) { *n = (size_t)fg->n; return 0; }
  size_t k = *n < (size_t)fg->n ? *n : (size_t)fg->n;
  for (size_t i = 0; i < k; i++) nodes[i] = &fg->nodes[i];
  *n = k;
  return 0;
}
EXPORT CUresult cuGraphNodeGetType(void *node, int *type) { (void)node; *type = 0; return 0; }
EXPORT CUresult cuGraphKernelNodeGetParams(void *node, knode_params_t *out) { *out = *(knode_params_t *)node; return 0; }
EXPORT CUresult cuGraphInstantiateWithFlags(void **exec, void *g, unsigned long long flags) {
  (void)flags;
  fake_exec_t *e = (fake_exec_t *)calloc(1, sizeof *e);
  if (!e) return 2;
  e->g = *(fake_graph_t *)g;
  *exec = e;
  return 0;
}
EXPORT CUresult cuGraphExecDestroy(void *exec) { free(exec); return 0; }
EXPORT CUresult cuGraphLaunch(void *exec, void *stream) {
  (void)stream;
  if (!t_has_ctx) return 201;
  fake_exec_t *e = (fake_exec_t *)exec;
  for (int i = 0; i < e->g.n; i++) note_launch();
  return 0;
}

static const struct { const char *name; void *fn; } g_self_table[] = {
  {"cuInit", (void *)cuInit},
  {"cuDriverGetVersion", (void *)cuDriverGetVersion},
  {"cuDeviceGetCount", (void *)cuDeviceGetCount},
  {"cuDeviceGet", (void *)cuDeviceGet},
  {"cuDeviceGetAttribute", (void *)cuDeviceGetAttribute},
  {"cuDeviceGetUuid", (void *)cuDeviceGetUuid},
  {"cuDeviceGetUuid_v2", (void *)cuDeviceGetUuid_v2},
  {"cuDeviceGetName", (void *)cuDeviceGetName},
  {"cuDeviceTotalMem_v2", (void *)cuDeviceTotalMem_v2},
  {"cuGetErrorString", (void *)cuGetErrorString},
  {"cuGetErrorName", (void *)cuGetErrorName},
  {"cuDevicePrimaryCtxRetain", (void *)cuDevicePrimaryCtxRetain},
  {"cuDevicePrimaryCtxRelease_v2", (void *)cuDevicePrimaryCtxRelease_v2},
  {"cuDevicePrimaryCtxRelease", (void *)cuDevicePrimaryCtxRelease_v2},
  {"cuDevicePrimaryCtxReset_v2", (void *)cuDevicePrimaryCtxReset_v2}, {"cuDevicePrimaryCtxReset", (void *)cuDevicePrimaryCtxReset_v2},
  {"cuDevicePrimaryCtxGetState", (void *)cuDevicePrimaryCtxGetState},
  {"cuCtxCreate_v2", (void *)cuCtxCreate_v2},
  {"cuCtxDestroy_v2", (void *)cuCtxDestroy_v2},
  {"cuCtxSetCurrent", (void *)cuCtxSetCurrent},
  {"cuCtxGetCurrent", (void *)cuCtxGetCurrent},
  {"cuCtxPushCurrent_v2", (void *)cuCtxPushCurrent_v2},
  {"cuCtxPopCurrent_v2", (void *)cuCtxPopCurrent_v2},
  {"cuCtxGetDevice", (void *)cuCtxGetDevice},
  {"cuCtxSynchronize", (void *)cuCtxSynchronize},
  {"cuCtxGetStreamPriorityRange", (void *)cuCtxGetStreamPriorityRange},
  {"cuMemAlloc_v2", (void *)cuMemAlloc_v2},
  {"cuMemAllocManaged", (void *)cuMemAllocManaged},
  {"cuMemAllocPitch_v2", (void *)cuMemAllocPitch_v2},
  {"cuMemAllocAsync", (void *)cuMemAllocAsync},
  {"cuMemAllocAsync_ptsz", (void *)cuMemAllocAsync_ptsz},
  {"cuMemAllocFromPoolAsync", (void *)cuMemAllocFromPoolAsync},
  {"cuMemAllocFromPoolAsync_ptsz", (void *)cuMemAllocFromPoolAsync_ptsz},
  {"cuMemFree_v2", (void *)cuMemFree_v2},
  {"cuMemFreeAsync", (void *)cuMemFreeAsync},
  {"cuMemFreeAsync_ptsz", (void *)cuMemFreeAsync_ptsz},
  {"cuMemGetInfo_v2", (void *)cuMemGetInfo_v2},
  {"cuMemCreate", (void *)cuMemCreate},
  {"cuMemRelease", (void *)cuMemRelease},
  {"cuArrayCreate_v2", (void *)cuArrayCreate_v2},
  {"cuArray3DCreate_v2", (void *)cuArray3DCreate_v2},
  {"cuMipmappedArrayCreate", (void *)cuMipmappedArrayCreate},
  {"cuArrayDestroy", (void *)cuArrayDestroy},
  {"cuMipmappedArrayDestroy", (void *)cuMipmappedArrayDestroy},
  {"cuMemHostAlloc", (void *)cuMemHostAlloc},
  {"cuMemFreeHost", (void *)cuMemFreeHost},
  {"cuMemHostGetDevicePointer_v2", (void *)cuMemHostGetDevicePointer_v2},
  {"cuMemsetD8_v2", (void *)cuMemsetD8_v2},
  {"cuMemcpyDtoH_v2", (void *)cuMemcpyDtoH_v2},
  {"cuMemcpyDtoHAsync_v2", (void *)cuMemcpyDtoHAsync_v2},
  {"cuMemcpyHtoD_v2", (void *)cuMemcpyHtoD_v2},
  {"cuMemcpyDtoD_v2", (void *)cuMemcpyDtoD_v2},
  {"cuMemcpy", (void *)cuMemcpy},
  {"cuStreamCreate", (void *)cuStreamCreate},
  {"cuStreamCreateWithPriority", (void *)cuStreamCreateWithPriority},
  {"cuStreamDestroy_v2", (void *)cuStreamDestroy_v2},
  {"cuStreamSynchronize", (void *)cuStreamSynchronize},
  {"cuStreamQuery", (void *)cuStreamQuery},
  {"cuStreamIsCapturing", (void *)cuStreamIsCapturing},
  {"cuStreamWaitValue64_v2", (void *)cuStreamWaitValue64_v2},
  {"cuStreamWaitValue64_v2_ptsz", (void *)cuStreamWaitValue64_v2_ptsz},
  {"cuStreamWriteValue64_v2", (void *)cuStreamWriteValue64_v2},
  {"cuStreamWriteValue64_v2_ptsz", (void *)cuStreamWriteValue64_v2_ptsz},
  {"cuModuleLoadData", (void *)cuModuleLoadData},
  {"cuModuleUnload", (void *)cuModuleUnload},
  {"cuModuleGetFunction", (void *)cuModuleGetFunction},
  {"cuFuncSetAttribute", (void *)cuFuncSetAttribute},
  {"cuFuncSetBlockShape", (void *)cuFuncSetBlockShape},
  {"cuLaunchKernel", (void *)cuLaunchKernel},
  {"cuLaunchKernel_ptsz", (void *)cuLaunchKernel_ptsz},
  {"cuLaunchKernelEx", (void *)cuLaunchKernelEx},
  {"cuGraphCreate", (void *)cuGraphCreate}, {"cuGraphDestroy", (void *)cuGraphDestroy},
  {"cuGraphAddKernelNode", (void *)cuGraphAddKernelNode}, {"cuGraphGetNodes", (void *)cuGraphGetNodes},
  {"cuGraphNodeGetType", (void *)cuGraphNodeGetType}, {"cuGraphKernelNodeGetParams", (void *)cuGraphKernelNodeGetParams},
  {"cuGraphInstantiateWithFlags", (void *)cuGraphInstantiateWithFlags}, {"cuGraphInstantiate", (void *)cuGraphInstantiateWithFlags},
  {"cuGraphExecDestroy", (void *)cuGraphExecDestroy}, {"cuGraphLaunch", (void *)cuGraphLaunch},
  {"cuLaunchKernelEx_ptsz", (void *)cuLaunchKernelEx_ptsz},
  {"cuLaunchCooperativeKernel", (void *)cuLaunchCooperativeKernel},
  {"cuLaunchCooperativeKernel_ptsz", (void *)cuLaunchCooperativeKernel_ptsz},
  {"cuLaunch", (void *)cuLaunch},
  {"cuLaunchGrid", (void *)cuLaunchGrid},
  {"cuLaunchGridAsync", (void *)cuLaunchGridAsync},
};
static void *self_lookup(const char *name) {
  for (size_t i = 0; i < sizeof g_self_table / sizeof g_self_table[0]; i++)
    if (!strcmp(g_self_table[i].name, name)) return g_self_table[i].fn;
  return NULL;
}
EXPORT CUresult cuGetProcAddress_v2(const char *sym, void **pfn, int ver, unsigned long long flags, void *status);
EXPORT CUresult cuGetProcAddress(const char *sym, void **pfn, int ver, unsigned long long flags);
static CUresult get_proc_address(const char *sym, void **pfn, unsigned long long flags) {
  char name[128];
  void *p = NULL;
  if (!strcmp(sym, "cuGetProcAddress")) p = (void *)cuGetProcAddress_v2;
  if (!p && (flags & 2)) { snprintf(name, sizeof name, "%s_ptsz", sym); p = self_lookup(name); }
  if (!p) { snprintf(name, sizeof name, "%s_v2", sym); p = self_lookup(name); }
  if (!p) p = self_lookup(sym);
  *pfn = p;
  return p ? 0 : 500;
}
/* both entry points use the static helper: an exported one calling the other would go through the PLT and land in a
 * preloaded library's hook, which the real driver's internal calls never do */
EXPORT CUresult cuGetProcAddress_v2(const char *sym, void **pfn, int ver, unsigned long long flags, void *status) {
  (void)ver; (void)status;
  return get_proc_address(sym, pfn, flags);
}
EXPORT CUresult cuGetProcAddress(const char *sym, void **pfn, int ver, unsigned long long flags) {
  (void)ver;
  return get_proc_address(sym, pfn, flags);
}

/* ------------------------------------------------------------------ NVML */
EXPORT nvmlReturn_t nvmlInit(void) { stub_init(); return 0; }
EXPORT nvmlReturn_t nvmlInit_v2(void) { stub_init(); return 0; }
EXPORT nvmlReturn_t nvmlInitWithFlags(unsigned f) { (void)f; stub_init(); return 0; }
EXPORT nvmlReturn_t nvmlShutdown(void) { return 0; }
EXPORT const char *nvmlErrorString(nvmlReturn_t r) { return r == 0 ? "Success" : r == 6 ? "Not Found" : "stub nvml error"; }
EXPORT nvmlReturn_t nvmlDeviceGetCount(unsigned *n) { stub_init(); *n = (unsigned)g_gpu_count; return 0; }
EXPORT nvmlReturn_t nvmlDeviceGetCount_v2(unsigned *n) { return nvmlDeviceGetCount(n); }
EXPORT nvmlReturn_t nvmlDeviceGetHandleByIndex(unsigned i, void **d) { stub_init(); if ((int)i >= g_gpu_count) return 2; *d = (void *)(uintptr_t)(0x9000 + i); return 0; }
EXPORT nvmlReturn_t nvmlDeviceGetHandleByIndex_v2(unsigned i, void **d) { return nvmlDeviceGetHandleByIndex(i, d); }
static int nv_idx(void *d) { return (int)((uintptr_t)d - 0x9000); }
EXPORT nvmlReturn_t nvmlDeviceGetIndex(void *d, unsigned *i) { int k = nv_idx(d); if (k < 0 || k >= g_gpu_count) return 2; *i = (unsigned)k; return 0; }
EXPORT nvmlReturn_t nvmlDeviceGetUUID(void *d, char *uuid, unsigned len) {
  unsigned char b[16];
  uuid_bytes(nv_idx(d), b);
  snprintf(uuid, len, "GPU-%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", b[0], b[1], b[2], b[3],
           b[4], b[5], b[6], b[7], b[8], b[9], b[10], b[11], b[12], b[13], b[14], b[15]);
  return 0;
}
static nvmlReturn_t proc_list(void *d, unsigned *count, vgpu_proc_t *out, int graphics) {
  int dev = nv_idx(d);
  unsigned n = 0, cap = *count;
  if (!graphics && g_any_ctx) {
    if (n < cap) { out[n].pid = (uint32_t)getpid(); out[n]._pad = 0; out[n].used_bytes = g_ctx_bytes + g_dev_bytes[dev & 15]; }
    n++;
  }
  for (int i = 0; i < g_nothers; i++) {
    if (graphics ? !g_others[i].graphics : !g_others[i].compute) continue;
    if (n < cap) { out[n].pid = g_others[i].pid; out[n]._pad = 0; out[n].used_bytes = g_others[i].bytes; }
    n++;
  }
  *count = n;
  return n > cap ? 7 /* INSUFFICIENT_SIZE */ : 0;
}
/* the 24-byte v3 ABI (pid, usedGpuMemory, gpuInstanceId, computeInstanceId), what a current NVML exports */
static nvmlReturn_t proc_list_v3(void *d, unsigned *count, vgpu_proc_v2_t *out, int graphics) {
  static __thread vgpu_proc_t narrow[VGPU_MAX_PIDS];
  unsigned cap = *count, n = cap > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : cap;
  nvmlReturn_t r = proc_list(d, &n, narrow, graphics);
  for (unsigned i = 0; i < n && i < cap; i++) {
    out[i].pid = narrow[i].pid; out[i]._pad = 0; out[i].used_bytes = narrow[i].used_bytes;
    out[i].gi = out[i].ci = 0xFFFFFFFFu;
  }
  *count = n;
  return r;
}
EXPORT nvmlReturn_t nvmlDeviceGetComputeRunningProcesses_v3(void *d, unsigned *c, vgpu_proc_v2_t *o) { return proc_list_v3(d, c, o, 0); }
EXPORT nvmlReturn_t nvmlDeviceGetGraphicsRunningProcesses_v3(void *d, unsigned *c, vgpu_proc_v2_t *o) { return proc_list_v3(d, c, o, 1); }
EXPORT nvmlReturn_t nvmlDeviceGetComputeRunningProcesses(void *d, unsigned *c, vgpu_proc_t *o) { return proc_list(d, c, o, 0); }
EXPORT nvmlReturn_t nvmlDeviceGetGraphicsRunningProcesses(void *d, unsigned *c, vgpu_proc_t *o) { return proc_list(d, c, o, 1); }
EXPORT nvmlReturn_t nvmlDeviceGetProcessUtilization(void *d, vgpu_util_sample_t *s, unsigned *count, unsigned long long since) {
  (void)d; (void)since;
  unsigned n = 0, cap = *count;
  uint64_t ts = now_us_wall();
  if (g_any_ctx) {
    if (n < cap) { memset(&s[n], 0, sizeof s[n]); s[n].pid = (uint32_t)getpid(); s[n].ts_us = ts; s[n].sm = (uint32_t)current_util(); }
    n++;
  }
  for (int i = 0; i < g_nothers; i++) {
    if (!g_others[i].compute || !g_others[i].sm) continue;
    if (n < cap) { memset(&s[n], 0, sizeof s[n]); s[n].pid = g_others[i].pid; s[n].ts_us = ts; s[n].sm = (uint32_t)g_others[i].sm; }
    n++;
  }
  *count = n;
  return n == 0 ? 6 : 0;
}
typedef struct { unsigned long long total, free, used; } nvmem_t;
typedef struct { unsigned version; unsigned long long total, reserved, free, used; } nvmem2_t;
static uint64_t dev_used(int dev) {
  uint64_t u = g_dev_bytes[dev & 15] + (g_any_ctx ? g_ctx_bytes : 0);
  for (int i = 0; i < g_nothers; i++) u += g_others[i].bytes;
  return u;
}
EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo(void *d, nvmem_t *m) {
  uint64_t u = dev_used(nv_idx(d));
  m->total = g_total_mem; m->used = u; m->free = g_total_mem - u;
  return 0;
}
EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo_v2(void *d, nvmem2_t *m) {
  uint64_t u = dev_used(nv_idx(d));
  m->total = g_total_mem; m->reserved = 512ull << 20; m->used = u; m->free = g_total_mem - u - m->reserved;
  return 0;
}
EXPORT nvmlReturn_t nvmlDeviceGetUtilizationRates(void *d, unsigned *u) { (void)d; u[0] = (unsigned)current_util(); u[1] = 0; return 0; }
EXPORT nvmlReturn_t nvmlDeviceSetComputeMode(void *d, int m) { (void)d; (void)m; return 0; }
EXPORT nvmlReturn_t nvmlDeviceGetPersistenceMode(void *d, int *m) { (void)d; *m = 1; return 0; }
