/*
 * storm.c - launch-storm tenant (BASELINE.json configs 0/1/4): N x cuLaunchKernel(grid 1,1,1)
 * of an empty kernel through the dynamic linker, per-call latency histogram, JSON on stdout.
 * TEST / BENCH INFRASTRUCTURE.  Works against the stub driver (no GPU: f == NULL) and against
 * the real driver (loads an empty kernel from PTX).
 *
 *   storm [--n N] [--threads T] [--sync-every K] [--device D] [--no-kernel]
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef int CUresult;
typedef CUresult (*launch_fn)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **);

static const char *k_ptx =
    ".version 7.0\n.target sm_52\n.address_size 64\n"
    ".visible .entry empty_kernel()\n{\n  ret;\n}\n";

static void *h_cuda;
static launch_fn p_launch;
static CUresult (*p_sync)(void);
static CUresult (*p_setctx)(void *);
static void *g_ctx, *g_func;
static long g_n = 1000000, g_sync_every = 0;
static int g_threads = 1;

static inline uint64_t now_ns(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

typedef struct { uint32_t *lat; long n; long fails; uint64_t t0, t1; } worker_t;

static void *worker(void *arg) {
  worker_t *w = (worker_t *)arg;
  p_setctx(g_ctx);
  w->t0 = now_ns();
  for (long i = 0; i < w->n; i++) {
    uint64_t a = now_ns();
    CUresult r = p_launch(g_func, 1, 1, 1, 1, 1, 1, 0, NULL, NULL, NULL);
    uint64_t b = now_ns();
    w->lat[i] = (uint32_t)((b - a) > 0xffffffffull ? 0xffffffffull : (b - a));
    w->fails += r != 0;
    if (g_sync_every && ((i + 1) % g_sync_every) == 0) p_sync();
  }
  w->t1 = now_ns();
  return NULL;
}

static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return (x > y) - (x < y);
}

int main(int argc, char **argv) {
  int device = 0, no_kernel = 0;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--n") && i + 1 < argc) g_n = atol(argv[++i]);
    else if (!strcmp(argv[i], "--threads") && i + 1 < argc) g_threads = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--sync-every") && i + 1 < argc) g_sync_every = atol(argv[++i]);
    else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--no-kernel")) no_kernel = 1;
  }
  h_cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h_cuda) { fprintf(stderr, "storm: %s\n", dlerror()); return 2; }
  CUresult (*p_init)(unsigned) = dlsym(h_cuda, "cuInit");
  CUresult (*p_get)(int *, int) = dlsym(h_cuda, "cuDeviceGet");
  CUresult (*p_retain)(void **, int) = dlsym(h_cuda, "cuDevicePrimaryCtxRetain");
  CUresult (*p_modload)(void **, const void *) = dlsym(h_cuda, "cuModuleLoadData");
  CUresult (*p_getfn)(void **, void *, const char *) = dlsym(h_cuda, "cuModuleGetFunction");
  p_setctx = dlsym(h_cuda, "cuCtxSetCurrent");
  p_sync = dlsym(h_cuda, "cuCtxSynchronize");
  p_launch = (launch_fn)dlsym(h_cuda, "cuLaunchKernel");
  int dev = 0;
  uint64_t t_init0 = now_ns();
  if (p_init(0) || p_get(&dev, device) || p_retain(&g_ctx, dev) || p_setctx(g_ctx)) { fprintf(stderr, "storm: init failed\n"); return 3; }
  if (!no_kernel) {
    void *mod = NULL;
    if (p_modload(&mod, k_ptx) || p_getfn(&g_func, mod, "empty_kernel")) { fprintf(stderr, "storm: module load failed\n"); return 4; }
  }
  /* warm-up: first launches pay lazy loading and, under a preload library, its bring-up */
  for (int i = 0; i < 2000; i++) p_launch(g_func, 1, 1, 1, 1, 1, 1, 0, NULL, NULL, NULL);
  p_sync();
  uint64_t t_init1 = now_ns();

  long per = g_n / g_threads;
  worker_t *ws = calloc((size_t)g_threads, sizeof *ws);
  pthread_t *th = calloc((size_t)g_threads, sizeof *th);
  for (int t = 0; t < g_threads; t++) { ws[t].n = per; ws[t].lat = malloc(sizeof(uint32_t) * (size_t)per); }
  uint64_t t0 = now_ns();
  for (int t = 1; t < g_threads; t++) pthread_create(&th[t], NULL, worker, &ws[t]);
  worker(&ws[0]);
  for (int t = 1; t < g_threads; t++) pthread_join(th[t], NULL);
  uint64_t t_host = now_ns();
  p_sync();
  uint64_t t1 = now_ns();

  long total = per * g_threads, fails = 0;
  uint32_t *all = malloc(sizeof(uint32_t) * (size_t)total);
  for (int t = 0; t < g_threads; t++) { memcpy(all + (size_t)t * per, ws[t].lat, sizeof(uint32_t) * (size_t)per); fails += ws[t].fails; }
  qsort(all, (size_t)total, sizeof(uint32_t), cmp_u32);
  double wall = (t1 - t0) * 1e-9, host = (t_host - t0) * 1e-9;
  unsigned long long sum = 0;
  for (long i = 0; i < total; i++) sum += all[i];
  printf("{\"launches\": %ld, \"threads\": %d, \"wall_s\": %.6f, \"host_s\": %.6f, \"launches_per_s\": %.1f, "
         "\"host_launches_per_s\": %.1f, \"p50_ns\": %u, \"p90_ns\": %u, \"p99_ns\": %u, \"p999_ns\": %u, \"max_ns\": %u, "
         "\"mean_ns\": %.1f, \"fails\": %ld, \"init_s\": %.4f}\n",
         total, g_threads, wall, host, total / wall, total / host, all[total / 2], all[(long)(total * 0.9)],
         all[(long)(total * 0.99)], all[(long)(total * 0.999)], all[total - 1], (double)sum / total, fails,
         (t_init1 - t_init0) * 1e-9);
  return 0;
}
