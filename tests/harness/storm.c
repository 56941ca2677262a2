/*
 * storm.c - launch-storm tenant (BASELINE.json configs 0/1/4): steps x launches-per-step
 * cuLaunchKernel(grid 1,1,1 / block 1,1,1) of an empty kernel through the dynamic linker, the
 * way an application linked against libcuda would issue them.  Per-call host latency
 * histogram, per-step device time (CUDA events on the launching stream), JSON on stdout.
 * TEST / BENCH INFRASTRUCTURE.  Works against the stub driver (no GPU: f == NULL) and the real
 * driver (empty kernel JIT-compiled from PTX).
 *
 *   storm [--steps K] [--warmup W] [--per-step L] [--threads T] [--sync-every S] [--device D]
 *         [--no-kernel] [--max-seconds X] [--n N (== --steps 1 --per-step N)]
 *         [--spin-iters I --grid G --block B]   busy kernel instead of the empty one
 *         [--block-with sync|htod|dtoh|dtod|copy]  which blocking call --sync-every issues (default cuCtxSynchronize;
 *                                               the others are 256 KiB synchronous copies on a scratch allocation)
 *         [--stream null|created|ptsz]          legacy stream (default), one created stream per thread, or the
 *                                               per-thread default stream through the _ptsz entry point
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
    ".visible .entry empty_kernel()\n{\n  ret;\n}\n"
    /* busy kernel for the fairness / utilisation cases: `iters` dependent adds per thread */
    ".visible .entry spin_kernel(.param .u32 iters)\n{\n"
    "  .reg .u32 %r<4>;\n  .reg .pred %p;\n"
    "  ld.param.u32 %r1, [iters];\n  mov.u32 %r2, 0;\n"
    "L0:\n  add.u32 %r2, %r2, 1;\n  setp.lt.u32 %p, %r2, %r1;\n  @%p bra L0;\n  ret;\n}\n";

static void *h_cuda;
static launch_fn p_launch;
static CUresult (*p_sync)(void);
static CUresult (*p_setctx)(void *);
static CUresult (*p_evcreate)(void **, unsigned);
static CUresult (*p_evrecord)(void *, void *);
static CUresult (*p_evsync)(void *);
static CUresult (*p_evelapsed)(float *, void *, void *);
static void *g_ctx, *g_func;
static int g_stream_mode; /* 0 legacy NULL stream, 1 one created stream per thread, 2 per-thread default stream (_ptsz entry) */
static CUresult (*p_stream_create)(void **, unsigned);
static long g_per_step = 200000, g_sync_every = 0;
static int g_block_with; /* 0 cuCtxSynchronize, 1 HtoD, 2 DtoH, 3 DtoD, 4 cuMemcpy */
static unsigned long long g_scratch;
static CUresult (*p_htod)(unsigned long long, const void *, size_t);
static CUresult (*p_dtoh)(void *, unsigned long long, size_t);
static CUresult (*p_dtod)(unsigned long long, unsigned long long, size_t);
static CUresult (*p_copy)(unsigned long long, unsigned long long, size_t);

#define BLOCK_BYTES (256u << 10) /* pageable and larger than the driver's inline-copy threshold: the call has to wait for the stream */
static CUresult block_now(void) {
  static __thread unsigned char *buf;
  if (!buf) buf = calloc(1, BLOCK_BYTES);
  switch (g_scratch ? g_block_with : 0) {
  case 1: return p_htod(g_scratch, buf, BLOCK_BYTES);
  case 2: return p_dtoh(buf, g_scratch, BLOCK_BYTES);
  case 3: return p_dtod(g_scratch + BLOCK_BYTES, g_scratch, BLOCK_BYTES);
  case 4: return p_copy(g_scratch, (unsigned long long)(uintptr_t)buf, BLOCK_BYTES); /* unified addressing: host -> device */
  default: return p_sync();
  }
}
static unsigned g_spin_iters = 0, g_grid = 1, g_block = 1;
static void *g_kparams[1];
static int g_threads = 1, g_steps = 1, g_warmup = 0;
static double g_max_seconds = 0;
static volatile int g_stop;
static uint64_t g_deadline_ns;

static inline uint64_t now_ns(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

typedef struct { uint32_t *lat; long n, done, fails; int record; } worker_t;

static void *worker(void *arg) {
  worker_t *w = (worker_t *)arg;
  p_setctx(g_ctx);
  void *stream = NULL;
  if (g_stream_mode == 1 && p_stream_create) p_stream_create(&stream, 1 /* CU_STREAM_NON_BLOCKING */);
  long i;
  for (i = 0; i < w->n && !g_stop; i++) {
    uint64_t a = now_ns();
    CUresult r = p_launch(g_func, g_grid, 1, 1, g_block, 1, 1, 0, stream, g_spin_iters ? g_kparams : NULL, NULL);
    uint64_t b = now_ns();
    if (w->record) w->lat[i] = (uint32_t)((b - a) > 0xffffffffull ? 0xffffffffull : (b - a));
    w->fails += r != 0;
    if (g_sync_every && ((i + 1) % g_sync_every) == 0) w->fails += block_now() != 0;
    if (g_deadline_ns && (i & 1023) == 0 && b > g_deadline_ns) g_stop = 1;
  }
  w->done = i;
  return NULL;
}

static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return (x > y) - (x < y);
}

int main(int argc, char **argv) {
  int device = 0, no_kernel = 0, host_index = -1;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--n") && i + 1 < argc) { g_per_step = atol(argv[++i]); g_steps = 1; }
    else if (!strcmp(argv[i], "--steps") && i + 1 < argc) g_steps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) g_warmup = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--per-step") && i + 1 < argc) g_per_step = atol(argv[++i]);
    else if (!strcmp(argv[i], "--threads") && i + 1 < argc) g_threads = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--sync-every") && i + 1 < argc) g_sync_every = atol(argv[++i]);
    else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--max-seconds") && i + 1 < argc) g_max_seconds = atof(argv[++i]);
    else if (!strcmp(argv[i], "--host-index") && i + 1 < argc) host_index = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--spin-iters") && i + 1 < argc) g_spin_iters = (unsigned)atol(argv[++i]);
    else if (!strcmp(argv[i], "--grid") && i + 1 < argc) g_grid = (unsigned)atol(argv[++i]);
    else if (!strcmp(argv[i], "--block") && i + 1 < argc) g_block = (unsigned)atol(argv[++i]);
    else if (!strcmp(argv[i], "--stream") && i + 1 < argc) { i++; g_stream_mode = !strcmp(argv[i], "created") ? 1 : !strcmp(argv[i], "ptsz") ? 2 : 0; }
    else if (!strcmp(argv[i], "--block-with") && i + 1 < argc) {
      i++;
      g_block_with = !strcmp(argv[i], "htod") ? 1 : !strcmp(argv[i], "dtoh") ? 2 : !strcmp(argv[i], "dtod") ? 3 : !strcmp(argv[i], "copy") ? 4 : 0;
    }
    else if (!strcmp(argv[i], "--no-kernel")) no_kernel = 1;
  }
  if (g_threads < 1) g_threads = 1;
  if (host_index < 0) host_index = device;
  h_cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h_cuda) { fprintf(stderr, "storm: %s\n", dlerror()); return 2; }
  CUresult (*p_init)(unsigned) = dlsym(h_cuda, "cuInit");
  CUresult (*p_get)(int *, int) = dlsym(h_cuda, "cuDeviceGet");
  CUresult (*p_retain)(void **, int) = dlsym(h_cuda, "cuDevicePrimaryCtxRetain");
  CUresult (*p_modload)(void **, const void *) = dlsym(h_cuda, "cuModuleLoadData");
  CUresult (*p_getfn)(void **, void *, const char *) = dlsym(h_cuda, "cuModuleGetFunction");
  p_setctx = dlsym(h_cuda, "cuCtxSetCurrent");
  p_sync = dlsym(h_cuda, "cuCtxSynchronize");
  p_launch = (launch_fn)dlsym(h_cuda, g_stream_mode == 2 ? "cuLaunchKernel_ptsz" : "cuLaunchKernel");
  p_stream_create = dlsym(h_cuda, "cuStreamCreate");
  p_evcreate = dlsym(h_cuda, "cuEventCreate");
  p_evrecord = dlsym(h_cuda, "cuEventRecord");
  p_evsync = dlsym(h_cuda, "cuEventSynchronize");
  p_evelapsed = dlsym(h_cuda, "cuEventElapsedTime");
  int have_events = p_evcreate && p_evrecord && p_evsync && p_evelapsed;
  int dev = 0;
  uint64_t t_init0 = now_ns();
  if (p_init(0) || p_get(&dev, device) || p_retain(&g_ctx, dev) || p_setctx(g_ctx)) { fprintf(stderr, "storm: init failed\n"); return 3; }
  if (!no_kernel) {
    void *mod = NULL;
    if (p_modload(&mod, k_ptx) || p_getfn(&g_func, mod, g_spin_iters ? "spin_kernel" : "empty_kernel")) { fprintf(stderr, "storm: module load failed\n"); return 4; }
    g_kparams[0] = &g_spin_iters;
  }
  if (g_block_with) {
    CUresult (*p_alloc)(unsigned long long *, size_t) = dlsym(h_cuda, "cuMemAlloc_v2");
    p_htod = dlsym(h_cuda, "cuMemcpyHtoD_v2");
    p_dtoh = dlsym(h_cuda, "cuMemcpyDtoH_v2");
    p_dtod = dlsym(h_cuda, "cuMemcpyDtoD_v2");
    p_copy = dlsym(h_cuda, "cuMemcpy");
    int ok = g_block_with == 1 ? !!p_htod : g_block_with == 2 ? !!p_dtoh : g_block_with == 3 ? !!p_dtod : !!p_copy;
    if (!ok || !p_alloc || p_alloc(&g_scratch, 2 * BLOCK_BYTES)) { fprintf(stderr, "storm: --block-with needs the copy entry point and a scratch allocation\n"); return 5; }
  }
  /* first launches pay lazy loading and, under a preload library, its device bring-up */
  for (int i = 0; i < (g_spin_iters ? 50 : 2000); i++) p_launch(g_func, g_grid, 1, 1, g_block, 1, 1, 0, NULL, g_spin_iters ? g_kparams : NULL, NULL);
  p_sync();
  uint64_t t_init1 = now_ns();

  long per_thread = g_per_step / g_threads;
  worker_t *ws = calloc((size_t)g_threads, sizeof *ws);
  pthread_t *th = calloc((size_t)g_threads, sizeof *th);
  size_t cap = (size_t)per_thread * (size_t)g_steps;
  uint32_t *all = malloc(sizeof(uint32_t) * cap * (size_t)g_threads);
  size_t nall = 0;
  for (int t = 0; t < g_threads; t++) ws[t].lat = malloc(sizeof(uint32_t) * (size_t)per_thread);
  void *ev0 = NULL, *ev1 = NULL;
  if (have_events && (p_evcreate(&ev0, 0) || p_evcreate(&ev1, 0))) have_events = 0;

  double *step_wall = calloc((size_t)g_steps, sizeof(double)), *step_dev = calloc((size_t)g_steps, sizeof(double));
  long total_done = 0, fails = 0;
  double timed_wall = 0, timed_host = 0;
  int steps_run = 0;
  for (int s = -g_warmup; s < g_steps && !g_stop; s++) {
    int timed = s >= 0;
    if (timed && s == 0 && g_max_seconds > 0) g_deadline_ns = now_ns() + (uint64_t)(g_max_seconds * 1e9);
    for (int t = 0; t < g_threads; t++) { ws[t].n = per_thread; ws[t].done = 0; ws[t].fails = 0; ws[t].record = timed; }
    p_sync();
    if (have_events) p_evrecord(ev0, NULL);
    uint64_t t0 = now_ns();
    for (int t = 1; t < g_threads; t++) pthread_create(&th[t], NULL, worker, &ws[t]);
    worker(&ws[0]);
    for (int t = 1; t < g_threads; t++) pthread_join(th[t], NULL);
    uint64_t t_host = now_ns();
    if (have_events) p_evrecord(ev1, NULL);
    p_sync();
    uint64_t t1 = now_ns();
    if (!timed) continue;
    float ms = 0;
    if (have_events && p_evsync(ev1) == 0 && p_evelapsed(&ms, ev0, ev1) == 0) step_dev[s] = ms * 1e-3;
    step_wall[s] = (t1 - t0) * 1e-9;
    timed_wall += step_wall[s];
    timed_host += (t_host - t0) * 1e-9;
    for (int t = 0; t < g_threads; t++) {
      memcpy(all + nall, ws[t].lat, sizeof(uint32_t) * (size_t)ws[t].done);
      nall += (size_t)ws[t].done;
      total_done += ws[t].done;
      fails += ws[t].fails;
    }
    steps_run++;
  }
  if (nall == 0) { printf("{\"launches\": 0}\n"); return 0; }
  qsort(all, nall, sizeof(uint32_t), cmp_u32);
  unsigned long long sum = 0;
  for (size_t i = 0; i < nall; i++) sum += all[i];
  unsigned long long (*metric)(int, int) = dlsym(RTLD_DEFAULT, "vgpu_b200_metric");
  /* limiter state of the B200 library, when it is the preloaded one (include/vgpu_b200.h) */
  struct { long long granted, consumed, bucket, share; int up_limit, sys_free, avg_sys_free, ctr_i, pre, valid,
           user_current, sys_current, sm_active_pct, queue_busy_pct; unsigned long long steps; } ls;
  memset(&ls, 0, sizeof ls);
  int (*lstate)(void *) = dlsym(RTLD_DEFAULT, "vgpu_b200_limiter_state");
  int have_ls = lstate && lstate(&ls) == 0;
  double dev_total = 0;
  for (int s = 0; s < steps_run; s++) dev_total += step_dev[s];
  printf("{\"launches\": %ld, \"steps\": %d, \"per_step\": %ld, \"threads\": %d, \"wall_s\": %.6f, \"host_s\": %.6f, "
         "\"device_s\": %.6f, \"launches_per_s\": %.1f, \"host_launches_per_s\": %.1f, \"p50_ns\": %u, \"p90_ns\": %u, "
         "\"p99_ns\": %u, \"p999_ns\": %u, \"max_ns\": %u, \"mean_ns\": %.1f, \"fails\": %ld, \"init_s\": %.4f, "
         "\"truncated\": %d, \"sampler_launches\": %llu, \"gated_launches\": %llu, \"watchdog_loans\": %llu, \"limiter\": {\"present\": %d, "
         "\"user_current\": %d, \"queue_busy_pct\": %d, \"sm_active_pct\": %d, \"share\": %lld, \"bucket\": %lld, "
         "\"control_steps\": %llu}, \"step_wall_s\": [",
         total_done, steps_run, g_per_step, g_threads, timed_wall, timed_host, dev_total, total_done / timed_wall,
         total_done / timed_host, all[nall / 2], all[(size_t)(nall * 0.9)], all[(size_t)(nall * 0.99)],
         all[(size_t)(nall * 0.999)], all[nall - 1], (double)sum / nall, fails, (t_init1 - t_init0) * 1e-9, (int)g_stop,
         metric ? metric(host_index, 7) : 0ull, metric ? metric(host_index, 0) : 0ull, metric ? metric(host_index, 9) : 0ull, have_ls, ls.user_current,
         ls.queue_busy_pct, ls.sm_active_pct, ls.share, ls.granted - ls.consumed, ls.steps);
  for (int s = 0; s < steps_run; s++) printf("%s%.6f", s ? ", " : "", step_wall[s]);
  printf("]}\n");
  return 0;
}
