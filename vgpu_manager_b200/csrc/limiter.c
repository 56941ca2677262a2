/*
 * limiter.c - compute-share limiter: launch hooks, the host side of the token bucket and the
 * tick thread that keeps the on-device sampler/controller running.
 *
 * Reference behaviour being replaced (library/src/cuda_hook.c):
 *   rate_limiter :308-330   a launch proceeds iff the bucket is >= 0, then pays
 *                           gridX*gridY*gridZ tokens; otherwise the CPU thread sleeps in 10 ms
 *                           steps until the watcher refills the bucket.
 *   utilization_watcher :380-471 + get_used_gpu_utilization :1044-1159   one CPU thread per <=4
 *                           devices polls NVML every ~80 ms and recomputes the refill share.
 *
 * B200 design:
 *   bucket   = granted - consumed.  `consumed` is host-owned (pinned page, one fetch_add per
 *              launch); `granted` is device-owned (HBM word + host-visible mirror), written only
 *              by the controller.  The hook reads the mirror with a single 8-byte load.
 *   gate     = when the bucket is empty the launch is NOT delayed on the CPU: a
 *              cuStreamWaitValue64(granted >= ticket) is enqueued in front of it, so the stream
 *              itself waits on the bucket word (see watchdog for which copy).  Contexts without
 *              64-bit stream mem-ops use vgpu_gate_kernel, a one-thread device spin.
 *   markers  = behind launches a cuStreamWriteValue64 bumps the stream's `done` sequence: the
 *              controller needs the first launch still parked behind the gate to know how many
 *              tokens were really spent, and the queue-busy signal compares it with `launched`.
 *   tick     = one light thread per process, the counterpart of the reference's watcher thread.
 *              Default (VGPU_B200_UTIL_SOURCE=nvml): once per control period (80 ms) it publishes
 *              the reference's own reading - the raw per-process NVML samples (or sm_util.config)
 *              plus the container-membership flags of their pids - into pinned memory and
 *              launches vgpu_refill_kernel (one warp in the common case), which folds them
 *              (cuda_hook.c:1044-1159) and runs the controller (:413-466) on the device.
 *              With the on-device signals (queue | sm | max) it launches vgpu_sampler_kernel every
 *              10 ms for a 0.5 ms window at a random offset instead (one CTA for `queue`, one per
 *              SM for the probe); the last CTA of every 8th launch runs the controller.  It
 *              also settles unmarked launch-train tails.
 *   blocking = driver calls that wait for the stream (cuCtxSynchronize, cuStreamSynchronize,
 *              cuEventSynchronize, the synchronous copies that involve host memory, cuMemFree)
 *              hold the context lock while they wait; behind a parked kernel they would keep the
 *              tick thread from launching the refill that ends them.  Their hooks therefore
 *              wait in user space - where the reference's thread would be asleep inside the launch
 *              hook - until nothing they depend on is parked (wait_until_unparked).
 *   watchdog = last resort for blocking calls that are not hooked: a second thread that never
 *              enters the driver.  The gate waits on the *host-visible mirror* of `granted`; when
 *              streams are parked, the controller has not stepped and the tick thread has been
 *              waiting inside the CUDA driver for two periods, the watchdog advances the mirror
 *              to the newest parked ticket (a loan: the controller folds it into `granted` at
 *              its next step, so it is repaid).
 *   governor = opt-in (VGPU_B200_GOVERNOR=1) alternative for GPUs that are not time-sliced with
 *              other contexts: vgpu_governor_kernel, one warp that stays resident while the
 *              tenant has work queued or parked, samples the queues every 50 us and runs the
 *              controller every 80 ms - no host involvement in the refill at all.  The launch
 *              hook (re)starts it; it retires when idle or ahead of a device-wide synchronise.
 */
#include "vgpu_internal.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>

extern vgpu_dev_rt *vgpu_rt_get(int host_index, CUdevice dev);

/* ------------------------------------------------------------------ stream slots
 * Every (stream, per-thread-default flag) of a device gets a slot in the pinned block: launch
 * sequence, completion marker, ticket ring.  Host-only bookkeeping lives here. */
#define MARK_EVERY 640u             /* dense launch trains: one completion marker per 640 launches.  Every marker is one more
                                       operation for the GPU front-end, which is what bounds an empty-kernel storm (one per
                                       256 cost 0.65 % of the launch rate against the reference).  Together with the
                                       run-ahead bound of VGPU_TICKET_RING - 64 = 960 outstanding launches the queue
                                       oscillates between ~320 and 960 entries: never empty (throughput stays GPU-bound),
                                       never full (the hook waits in user space once per 640 launches instead of every
                                       call blocking inside the driver: p50 1.4 us instead of 2.07 us, measured) */
#define SPARSE_TSC 120000ull        /* launches further apart than ~50 us are each marked (on-device signals only) */
#define SLOT_TOMB ((uintptr_t)1)    /* key of a slot whose stream was destroyed: reusable, but does not end a probe chain */
typedef struct {
  volatile uintptr_t key;           /* CUstream | ptsz bit | top bit; 0 = never used, SLOT_TOMB = freed */
  CUstream stream;
  int ptsz;
  volatile int lock;                /* serialises {ticket, sequence, ring write} of launches sharing the slot */
  volatile unsigned long long marked; /* last sequence number a marker was enqueued for */
  unsigned long long seen_launched;   /* tick thread: launch count at the previous tick  */
  unsigned long long last_tsc;
} slot_t;
static slot_t g_slots[VGPU_MAX_DEVICES][VGPU_STREAM_SLOTS];
static pthread_mutex_t g_slot_insert_mu = PTHREAD_MUTEX_INITIALIZER;

static inline unsigned long long rdtsc(void) {
  unsigned int lo, hi;
  __asm__ volatile("rdtsc" : "=a"(lo), "=d"(hi));
  return ((unsigned long long)hi << 32) | lo;
}

static inline void slot_lock(slot_t *sl) {
  while (__sync_lock_test_and_set(&sl->lock, 1))
    while (sl->lock) __builtin_ia32_pause();
}
static inline void slot_unlock(slot_t *sl) { __sync_lock_release(&sl->lock); }

static inline uintptr_t slot_key(CUstream s, int ptsz) {
  return ((uintptr_t)s << 1) | (uintptr_t)(ptsz & 1) | ((uintptr_t)1 << 63);
}

/* Open addressing with tombstones: a destroyed stream's slot keeps later probe chains intact
 * (a live stream that originally probed past it must still find its own slot).  Lookups are
 * lock-free; insertions (once per stream) are serialised and re-probe under the mutex, so one
 * stream first used from two threads at once cannot end up with two slots. */
static uint32_t slot_insert(slot_t *tab, uintptr_t key, uint32_t h, CUstream s, int ptsz) {
  uint32_t out = VGPU_STREAM_SLOTS - 1, reuse = VGPU_STREAM_SLOTS;
  pthread_mutex_lock(&g_slot_insert_mu);
  for (uint32_t i = 0; i < VGPU_STREAM_SLOTS - 1; i++) {
    uint32_t idx = (h + i) % (VGPU_STREAM_SLOTS - 1);
    uintptr_t k = tab[idx].key;
    if (k == key) { out = idx; reuse = VGPU_STREAM_SLOTS; break; }
    if (k == SLOT_TOMB && reuse == VGPU_STREAM_SLOTS) reuse = idx;
    if (k == 0) { if (reuse == VGPU_STREAM_SLOTS) reuse = idx; break; }
  }
  if (reuse != VGPU_STREAM_SLOTS) {
    tab[reuse].stream = s;
    tab[reuse].ptsz = ptsz;
    tab[reuse].marked = tab[reuse].seen_launched = 0;
    __sync_synchronize();
    tab[reuse].key = key;
    out = reuse;
  }
  pthread_mutex_unlock(&g_slot_insert_mu);
  return out; /* table full: the overflow slot, shared - every launch is marked */
}

static inline uint32_t slot_of(int host_index, CUstream s, int ptsz) {
  uintptr_t key = slot_key(s, ptsz);
  uint32_t h = (uint32_t)((key >> 4) * 0x9E3779B97F4A7C15ull >> 58); /* 6 bits */
  slot_t *tab = g_slots[host_index];
  for (uint32_t i = 0; i < VGPU_STREAM_SLOTS - 1; i++) {
    uint32_t idx = (h + i) % (VGPU_STREAM_SLOTS - 1);
    uintptr_t k = tab[idx].key;
    if (likely(k == key)) return idx;
    if (k == 0) break;
  }
  return slot_insert(tab, key, h, s, ptsz);
}

/* ------------------------------------------------------------------ tick thread */
static pthread_once_t g_tick_once = PTHREAD_ONCE_INIT;
static volatile unsigned g_tick_epoch;
static volatile int g_tick_devices[VGPU_MAX_DEVICES]; /* host indexes with a live runtime + core limit */
static uint32_t g_gov_interval_us = 50, g_gov_period_us = 80000, g_gov_idle_us = 5000;
static int g_governor_mode; /* VGPU_B200_GOVERNOR=1 */
static int g_skip_idle = 1;  /* VGPU_B200_SKIP_IDLE_WINDOWS=0 turns the skipping of idle sampler windows off */
static uint32_t g_window_us = 500, g_interval_us = 50, g_period_ticks = 8, g_tick_ms = 10;
static volatile int g_sync_waiters; /* threads currently inside a device-wide synchronise */
static volatile unsigned long g_tick_gen, g_wd_gen; /* completed loop iterations (quiescence points) */

static uint32_t env_u32(const char *name, uint32_t dflt) {
  const char *s = vgpu_tunable(name); /* ignored under a mounted control-plane config */
  return s ? (uint32_t)strtoul(s, NULL, 10) : dflt;
}

/* ------------------------------------------------------------------ utilisation publication (L5, host half)
 * What get_gpu_process_from_local_nvml_driver / _external_watcher collect (cuda_hook.c:922-1042),
 * left raw: the fold itself runs on the device.  The outcomes the reference distinguishes are
 * kept apart (status): list query failed -> nothing changes; sample query failed (NVML answers
 * NOT_FOUND between its ~1 s sample updates) -> only the process count changes and the watcher
 * keeps steering with the previous reading; both fine -> fold. */
static unsigned long long wall_us(void) {
  struct timespec now;
  clock_gettime(CLOCK_REALTIME, &now);
  return (unsigned long long)now.tv_sec * 1000000ull + (unsigned long long)now.tv_nsec / 1000ull;
}

static int publish_from_external_watcher(int h, vgpu_util_req_t *U) {
  if (!G_cfg->sm_watcher || !G_smutil) return 0;
  int fd = vgpu_smutil_rdlock(h);
  if (fd < 0) {
    VLOG(VL_WARNING, "failed to acquire read lock for host device %d, fallback to nvml driver", h);
    return 0;
  }
  const vgpu_smutil_dev_t *d = &G_smutil->devices[h];
  int ok = 0;
  if (wall_us() - d->last_seen_us < 5000000ull) { /* is_expired (:1002-1007) */
    uint32_t n = d->samples_size > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : d->samples_size;
    memcpy(U->samples, d->samples, (size_t)n * sizeof(vgpu_util_sample_t));
    U->n_samples = n;
    U->sys_process_num = (int)(d->compute_size >= d->graphics_size ? d->compute_size : d->graphics_size);
    U->checktime_us = d->last_seen_us;
    U->status = VGPU_UTIL_SAMPLES;
    ok = 1;
  }
  vgpu_smutil_unlock(fd, h);
  return ok;
}

static void publish_from_nvml(int h, vgpu_util_req_t *U) {
  nvmlDevice_t nv = vgpu_nvml_handle_of_host(h);
  if (!nv) return;
  static vgpu_proc_t procs[VGPU_MAX_PIDS];
  static vgpu_proc_v2_t wide[VGPU_MAX_PIDS];
  unsigned int n = VGPU_MAX_PIDS;
  nvmlReturn_t r;
  if (R.nvmlDeviceGetComputeRunningProcesses) r = R.nvmlDeviceGetComputeRunningProcesses(nv, &n, procs);
  else if (R.nvmlDeviceGetComputeRunningProcesses_v3) r = R.nvmlDeviceGetComputeRunningProcesses_v3(nv, &n, wide);
  else r = NVML_ERROR_FUNCTION_NOT_FOUND;
  if (r != NVML_SUCCESS) {
    VLOG(VL_VERBOSE, "nvmlDeviceGetComputeRunningProcesses can't get pids on host device %d, return %d, str: %s",
         h, r, vgpu_nv_err(r));
    return; /* VGPU_UTIL_NOTHING */
  }
  U->sys_process_num = (int)n;
  if (n == 0) { /* nobody computing: the graphics list decides the process count (:952-971) */
    n = VGPU_MAX_PIDS;
    if (R.nvmlDeviceGetGraphicsRunningProcesses) r = R.nvmlDeviceGetGraphicsRunningProcesses(nv, &n, procs);
    else if (R.nvmlDeviceGetGraphicsRunningProcesses_v3) r = R.nvmlDeviceGetGraphicsRunningProcesses_v3(nv, &n, wide);
    else r = NVML_ERROR_FUNCTION_NOT_FOUND;
    if (r == NVML_SUCCESS) U->sys_process_num = (int)n;
  }
  U->status = VGPU_UTIL_NPROC_ONLY;
  U->checktime_us = wall_us() - 1000000ull; /* "since one second ago" (:972-976) */
  unsigned int ns = VGPU_MAX_PIDS;
  r = R.nvmlDeviceGetProcessUtilization ? R.nvmlDeviceGetProcessUtilization(nv, U->samples, &ns, U->checktime_us)
                                        : NVML_ERROR_FUNCTION_NOT_FOUND;
  if (r != NVML_SUCCESS) {
    if (r != NVML_ERROR_NOT_FOUND)
      VLOG(VL_VERBOSE, "nvmlDeviceGetProcessUtilization can't get process utilization on host device %d, return: %d, str: %s",
           h, r, vgpu_nv_err(r));
    return;
  }
  U->n_samples = ns > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : ns;
  U->status = VGPU_UTIL_SAMPLES;
}

/* ------------------------------------------------------------------ node-level rebalance (host half)
 * Optional VGPU_CFG_DIR/rebalance.config (kernel_abi.h): the node agent's assignment for this GPU
 * is handed to the controller through the pinned block; while that file exists the tenant
 * publishes its state next to the lock file so the agent can fold it into its next plan. */
static void exchange_with_node_agent(vgpu_dev_rt *rt, int h) {
  static int fd_cfg = -2; /* -2 not looked for yet, -1 absent (looked for again every ~5 s) */
  static unsigned looked;
  if (fd_cfg < 0) {
    if (fd_cfg == -1 && (++looked % 64) != 0) return;
    fd_cfg = open(VP(VGPU_REBALANCE_FILE), O_RDONLY | O_CLOEXEC);
    if (fd_cfg < 0) { fd_cfg = -1; return; }
  }
  vgpu_rebalance_rec_t rec;
  if (pread(fd_cfg, &rec, sizeof rec, (off_t)h * (off_t)sizeof rec) != (ssize_t)sizeof rec || rec.magic != VGPU_REBALANCE_MAGIC) return;
  vgpu_lim_host_t *H = rt->lim_h;
  if (rec.seq != H->ext_limits_seq) {
    H->ext_up_limit = rec.up_limit;
    H->ext_soft_core = rec.soft_core;
    __sync_synchronize();
    H->ext_limits_seq = rec.seq;
    VLOG(VL_INFO, "host device %d: node agent assigned up_limit %d soft_core %d (seq %u)", h, rec.up_limit, rec.soft_core, rec.seq);
  }
  static int fd_st[VGPU_MAX_DEVICES];
  if (!fd_st[h]) {
    char raw[64];
    snprintf(raw, sizeof raw, VGPU_STATUS_FMT, h);
    int fd = open(VP(raw), O_RDWR | O_CREAT | O_CLOEXEC, 0644);
    fd_st[h] = fd >= 0 ? fd + 1 : -1;
  }
  if (fd_st[h] > 0) {
    static uint32_t seq;
    const vgpu_cfg_dev_t *c = &G_cfg->devices[h];
    vgpu_tenant_status_t st = {VGPU_REBALANCE_MAGIC, ++seq, (int32_t)getpid(), H->user_current, H->sys_current, H->up_limit_mirror,
                               c->hard_core, c->soft_core, H->share_mirror, vgpu_metric_get(h, VM_RATE_GATED),
                               vgpu_metric_get(h, VM_RATE_GATED) + vgpu_metric_get(h, VM_RATE_FAST), H->steps};
    ssize_t w = pwrite(fd_st[h] - 1, &st, sizeof st, 0);
    (void)w;
  }
}

/* ------------------------------------------------------------------ on-device readings for the node's SM watcher (8f-1)
 * A tenant whose limiter steers on an on-device signal (queue-busy / per-SM probe, or the governor) knows its own SM
 * utilisation without NVML.  Once per control period the reading goes into VGPU_LOCK_DIR/vgpu_<h>.readings
 * (include/vgpu_contract.h), from where vgpu-smwatcher --source device|mixed builds the per-process samples of
 * sm_util.config for every consumer on the node - the counterpart of watcher.go:160-182 without the
 * nvmlDeviceGetProcessUtilization poll.  Plain stores into a shared mapping: no lock, no driver call, no syscall
 * after the first publication. */
static void publish_device_reading(int h, const vgpu_lim_host_t *H) {
  static vgpu_readings_t *map[VGPU_MAX_DEVICES];
  static int slot[VGPU_MAX_DEVICES];    /* index + 1 */
  static unsigned epoch_of[VGPU_MAX_DEVICES];
  static uint64_t me;
  if (epoch_of[h] != vgpu_fork_epoch + 1) { /* a forked child is another owner */
    epoch_of[h] = vgpu_fork_epoch + 1;
    slot[h] = 0;
    me = 0;
  }
  if (map[h] == (vgpu_readings_t *)MAP_FAILED) return;
  if (!map[h]) {
    char raw[64];
    snprintf(raw, sizeof raw, VGPU_READINGS_FMT, h);
    int fd = open(VP(raw), O_RDWR | O_CREAT | O_CLOEXEC, 0666);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0 || ((size_t)st.st_size < sizeof(vgpu_readings_t) && ftruncate(fd, (off_t)sizeof(vgpu_readings_t)) != 0)) {
      if (fd >= 0) close(fd);
      map[h] = (vgpu_readings_t *)MAP_FAILED;
      return;
    }
    if (st.st_size == 0 && fchmod(fd, 0666) != 0) { /* created here: tenants of other uids publish into it too */ }
    void *m = mmap(NULL, sizeof(vgpu_readings_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    map[h] = (vgpu_readings_t *)m; /* MAP_FAILED: never tried again */
    if (m == MAP_FAILED) return;
  }
  if (!me) {
    struct stat ns;
    uint64_t ino = stat("/proc/self/ns/pid", &ns) == 0 ? (uint64_t)ns.st_ino : 0;
    me = (ino << 32) | (uint32_t)getpid();
  }
  vgpu_readings_t *F = map[h];
  const unsigned long long now = wall_us();
  if (!slot[h] || F->slots[slot[h] - 1].owner != me) { /* claim: my old slot, a free one, or one nobody wrote for 10 s */
    slot[h] = 0;
    for (int pass = 0; pass < 2 && !slot[h]; pass++)
      for (int i = 0; i < VGPU_READINGS_SLOTS; i++) {
        uint64_t o = F->slots[i].owner;
        if (o == me) { slot[h] = i + 1; break; }
        if (pass == 0) continue; /* first pass only looks for a slot that is already mine */
        if (o != 0 && now - F->slots[i].ts_us < 10000000ull) continue;
        if (__sync_bool_compare_and_swap(&F->slots[i].owner, o, me)) { F->slots[i].ts_us = now; F->slots[i].seq = 0; slot[h] = i + 1; break; }
      }
    if (!slot[h]) return; /* 1024 live publishers on one GPU: give up quietly */
  }
  vgpu_reading_t *r = &F->slots[slot[h] - 1];
  int u = H->user_current;
  r->sm_pct = (uint32_t)(u < 0 ? 0 : u > 100 ? 100 : u);
  r->queue_busy_pct = (uint32_t)H->queue_busy_pct;
  r->sm_active_pct = (uint32_t)H->sm_active_pct;
  r->seq++;
  __sync_synchronize();
  r->ts_us = now;
}

static unsigned refill_block(const vgpu_util_req_t *U) {
  unsigned n = U->status == VGPU_UTIL_SAMPLES ? U->n_samples : 0;
  return n ? (n + 31u) & ~31u : 32u;
}

/* returns the CTA size the refill kernel needs for this publication */
static unsigned publish_utilization(int h, vgpu_util_req_t *U) {
  U->status = VGPU_UTIL_NOTHING;
  U->n_samples = 0;
  U->mode = (uint32_t)G_cfg->compatibility_mode;
  U->have_container_pids = 1;
  if (!publish_from_external_watcher(h, U)) publish_from_nvml(h, U);
  if (U->status == VGPU_UTIL_SAMPLES && G_cfg->compatibility_mode != VGPU_MODE_HOST) {
    static uint32_t pids[VGPU_MAX_PIDS];
    for (uint32_t i = 0; i < U->n_samples; i++) pids[i] = U->samples[i].pid;
    U->have_container_pids = (uint32_t)vgpu_pid_flags_util(pids, U->n_samples, U->flags);
  }
  __sync_synchronize();
  U->seq++;
  return refill_block(U);
}

/* ------------------------------------------------------------------ the watcher starts at cuInit
 * The reference spawns its watcher at the first successful cuInit (cuda_hook.c:566-577): it steps
 * - and lets the bucket fill - while the application is still creating its context and loading
 * modules.  The controller here lives in the tenant's context, which does not exist yet.  So the
 * tick thread is started at cuInit too and, until the device runtime is up, keeps every
 * period's publication in a backlog; the moment the runtime appears the backlog is replayed through
 * vgpu_refill_kernel step by step, which leaves the device state exactly where the reference's
 * watcher would be.  It matters for more than the first burst: tenants that start together reach
 * their first launch 0.3 s apart (the driver serialises context creation), and a controller that only
 * starts then gives the first tenant a head start of several million tokens of `share`; the
 * increments are the same for everybody once they see the same reading, so that offset never
 * decays (measured: max/min 2-5 between four 25 % tenants, 1.0-1.5 under the reference). */
#define BACKLOG_MAX 64u /* ~5 s of control periods; beyond that the controller has long saturated */
typedef struct {
  vgpu_util_req_t *ring; /* BACKLOG_MAX publications, allocated at the first push */
  uint32_t head, count;
} backlog_t;
static backlog_t g_backlog[VGPU_MAX_DEVICES];

static void backlog_push(int h) {
  backlog_t *b = &g_backlog[h];
  if (!b->ring) b->ring = (vgpu_util_req_t *)calloc(BACKLOG_MAX, sizeof(vgpu_util_req_t));
  if (!b->ring) return;
  uint32_t at = (b->head + b->count) % BACKLOG_MAX;
  if (b->count == BACKLOG_MAX) b->head = (b->head + 1) % BACKLOG_MAX; /* full: the oldest goes */
  else b->count++;
  publish_utilization(h, &b->ring[at]);
}

/* caller has the runtime's context current */
static void backlog_replay(vgpu_dev_rt *rt, int h) {
  backlog_t *b = &g_backlog[h];
  for (uint32_t i = 0; i < b->count; i++) {
    const vgpu_util_req_t *src = &b->ring[(b->head + i) % BACKLOG_MAX];
    uint32_t seq = rt->u_req->seq + 1;
    memcpy(rt->u_req, src, offsetof(vgpu_util_req_t, samples) + (size_t)(src->status == VGPU_UTIL_SAMPLES ? src->n_samples : 0) * sizeof(vgpu_util_sample_t));
    memcpy(rt->u_req->flags, src->flags, src->status == VGPU_UTIL_SAMPLES ? src->n_samples : 0);
    rt->u_req->seq = seq;
    __sync_synchronize();
    unsigned block = refill_block(rt->u_req);
    void *params[] = {&rt->lim_d, &rt->lim_h_d, &rt->u_req_d};
    if (VGPU_CAPCHK(R.cuLaunchKernel(rt->k_refill, 1, 1, 1, block, 1, 1, 0, rt->s_stream, params, NULL)) != CUDA_SUCCESS) break;
    if (VGPU_CAPCHK(R.cuStreamSynchronize(rt->s_stream)) != CUDA_SUCCESS) break; /* the block is reused by the next step */
    vgpu_metric_add(h, VM_SAMPLER_LAUNCHES, 1);
  }
  VLOG(VL_INFO, "host device %d: replayed %u control periods that elapsed before the device runtime came up", h, b->count);
  b->count = b->head = 0;
  free(b->ring);
  b->ring = NULL;
}

/* External SM watcher (reference cuda_hook.c:1009-1042): when the control plane publishes
 * sm_util.config the per-process lists and utilisation samples of every tenant on the GPU are
 * read from that file (byte-range read lock, 5 s staleness) instead of from NVML. */
static int refresh_from_external_watcher(vgpu_dev_rt *rt) {
  if (!G_cfg->sm_watcher || !G_smutil) return 0;
  int h = rt->host_index;
  int fd = vgpu_smutil_rdlock(h);
  if (fd < 0) return 0;
  const vgpu_smutil_dev_t *d = &G_smutil->devices[h];
  struct timespec now;
  clock_gettime(CLOCK_REALTIME, &now);
  unsigned long long now_us = (unsigned long long)now.tv_sec * 1000000ull + (unsigned long long)now.tv_nsec / 1000ull;
  int ok = 0;
  if (now_us - d->last_seen_us < 5000000ull) {
    unsigned n = d->compute_size >= d->graphics_size ? d->compute_size : d->graphics_size;
    rt->lim_h->ext_sys_process_num = n ? (int)n : 1;
    int total = 0;
    unsigned ns = d->samples_size > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : d->samples_size;
    for (unsigned i = 0; i < ns; i++) {
      if (d->samples[i].ts_us < d->last_seen_us) continue;
      unsigned sm = d->samples[i].sm <= 100 ? d->samples[i].sm : 0;
      unsigned enc = d->samples[i].enc <= 100 ? d->samples[i].enc : 0, dec = d->samples[i].dec <= 100 ? d->samples[i].dec : 0;
      total += (int)(sm + (enc + dec) * 85 / 100);
    }
    int others = total - rt->lim_h->user_current;
    rt->lim_h->ext_sys_current = others > 0 ? others : 0;
    ok = 1;
  }
  vgpu_smutil_unlock(fd, h);
  return ok;
}

static void refresh_process_count(vgpu_dev_rt *rt) {
  /* sys_process_num feeds the jitter guard and the balance policy (cuda_hook.c:424,:431-449);
   * it changes on process start/exit, so one cheap list query per second is plenty */
  if (refresh_from_external_watcher(rt)) return;
  nvmlDevice_t nv = vgpu_nvml_handle_of_host(rt->host_index);
  if (!nv || !R.nvmlDeviceGetComputeRunningProcesses) return;
  static vgpu_proc_t procs[VGPU_MAX_PIDS];
  unsigned int n = VGPU_MAX_PIDS;
  if (R.nvmlDeviceGetComputeRunningProcesses(nv, &n, procs) == NVML_SUCCESS)
    rt->lim_h->ext_sys_process_num = n ? (int)n : 1;
  if (!G_cfg->devices[rt->host_index].hard_limit && R.nvmlDeviceGetUtilizationRates) {
    /* balance policy needs the whole GPU's load; other tenants are invisible from inside
     * this context, so take the device-level figure and let the kernel subtract ours */
    vnv_utilization_t u;
    if (R.nvmlDeviceGetUtilizationRates(nv, &u) == NVML_SUCCESS) {
      int others = (int)u.gpu - rt->lim_h->user_current;
      rt->lim_h->ext_sys_current = others > 0 ? others : 0;
    }
  }
}

/* A dense launch train only carries a marker every MARK_EVERY launches, so its tail may stay
 * unmarked.  If a stream has not launched since the previous tick and the driver says it is
 * idle, everything it was given has completed. */
/* the tick thread is between cuCtxPushCurrent and cuCtxPopCurrent and not in NVML (read by the watchdog) */
static volatile int g_tick_in_cuda;

/* LOGGER_LEVEL >= 4 only: name any step of the refill path, or any hooked blocking call, that took longer
 * than a quarter of a control period - the first question when a watchdog loan shows up in a log */
static int g_trace_slow = -1;
static inline uint64_t slow_t0(void) {
  if (unlikely(g_trace_slow < 0)) g_trace_slow = vgpu_log_level() >= VL_VERBOSE;
  if (likely(!g_trace_slow)) return 0;
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
static inline void slow_end(uint64_t t0, const char *what) {
  if (likely(!t0)) return;
  uint64_t dt = slow_t0() - t0;
  if (dt > 20000000ull) VLOG(VL_VERBOSE, "slow: %s took %llu ms", what, (unsigned long long)(dt / 1000000ull));
}

static void settle_idle_streams(vgpu_dev_rt *rt, int h) {
  vgpu_lim_host_t *H = rt->lim_h;
  for (uint32_t i = 0; i < VGPU_STREAM_SLOTS - 1; i++) {
    slot_t *sl = &g_slots[h][i];
    if (sl->key <= SLOT_TOMB) continue;
    unsigned long long l = H->launched[i], d = H->done[i];
    int quiet = (l == sl->seen_launched);
    sl->seen_launched = l;
    if (l <= d || !quiet || sl->marked >= l) continue;
    if (sl->ptsz && sl->stream == NULL) continue; /* another thread's default stream: not addressable */
    int capturing = 0; /* a query would invalidate the tenant's capture */
    if (sl->stream && R.cuStreamIsCapturing && (R.cuStreamIsCapturing(sl->stream, &capturing) != CUDA_SUCCESS || capturing)) continue;
    if (R.cuStreamQuery && VGPU_CAPCHK(R.cuStreamQuery(sl->stream)) == CUDA_SUCCESS && H->launched[i] == l && H->done[i] < l)
      H->done[i] = l;
  }
}

/* Decide whether this tick needs a sampler launch.  Same predicate as the kernel's queue-busy
 * sample, evaluated on the host copies: a stream is executing iff its oldest unfinished launch is
 * admitted.  Returns 1 = launch; 0 = skipped (and counted in *skipped). */
static uint32_t g_skipped[VGPU_MAX_DEVICES];
static int tenant_activity(vgpu_dev_rt *rt, uint32_t *skipped) {
  vgpu_lim_host_t *H = rt->lim_h;
  /* the SM probe also sees work the hook does not number (graph replays): it needs its windows */
  if (!g_skip_idle || H->util_source != VGPU_SRC_QUEUE) return 1;
  long long granted = H->granted_mirror;
  int parked = 0;
  for (uint32_t i = 0; i < VGPU_STREAM_SLOTS; i++) {
    unsigned long long d = H->done[i], l = H->launched[i];
    if (l <= d) continue;
    if (granted - H->ticket[i][(d + 1) & (VGPU_TICKET_RING - 1)] >= 0) return 1; /* executing */
    parked = 1;
  }
  if (H->release_pending) return 1;                       /* a loan is waiting to be folded in */
  if (parked && *skipped + 1 >= g_period_ticks) return 1; /* the period boundary: the refill that releases them */
  if (*skipped < 0x3fffffffu) (*skipped)++;
  return 0;
}

/* same reading of the knob as device.c read_util_tunables: anything but queue | sm | max is the default NVML reading */
static int source_is_nvml(void) {
  const char *s = vgpu_tunable("VGPU_B200_UTIL_SOURCE");
  return !s || (strcmp(s, "queue") && strcmp(s, "sm") && strcmp(s, "max"));
}

static void *tick_main(void *arg) {
  (void)arg;
  uint32_t epoch = 0;
  int fails = 0;
  uint64_t rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)getpid();
  int relaxed = 0;
  g_tick_in_cuda = 0; /* a forked child inherits the flag as the parent's tick thread left it */
  for (;;) {
    /* One short sampler window per tick, at a uniformly random offset inside the tick: the
     * sampler stays resident ~5 % of the time (it would otherwise show up as GPU utilisation in
     * NVML) and cannot phase-lock with the refill bursts the controller itself causes. */
    rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
    uint64_t tick_ns = (uint64_t)g_tick_ms * 1000000ull, win_ns = (uint64_t)g_window_us * 1000ull;
    uint64_t slack = tick_ns > win_ns ? tick_ns - win_ns : 0;
    uint64_t before = slack ? rng % slack : 0;
    struct timespec nap = {(time_t)(before / 1000000000ull), (long)(before % 1000000000ull)};
    nanosleep(&nap, NULL);
    if (g_tick_epoch != vgpu_fork_epoch + 1) return NULL;
    epoch++;
    for (int h = 0; h < VGPU_MAX_DEVICES; h++) {
      if (!g_tick_devices[h]) continue;
      vgpu_dev_rt *rt = vgpu_rt_peek(h);
      if (!rt) {
        /* no context yet: the watcher's readings are kept for the controller to catch up on */
        if (!g_governor_mode && epoch % g_period_ticks == 0 && G_cfg->devices[h].core_limit &&
            source_is_nvml()) backlog_push(h);
        continue;
      }
      if ((rt->lim_h->util_source != VGPU_SRC_NVML || g_governor_mode) && epoch % g_period_ticks == 0)
        publish_device_reading(h, rt->lim_h); /* for the node's SM watcher; the default NVML reading has nothing to add */
      uint64_t ts0 = slow_t0();
      g_tick_in_cuda = 1;
      if (VGPU_CAPCHK(R.cuCtxPushCurrent_v2(rt->ctx)) != CUDA_SUCCESS) { g_tick_in_cuda = 0; continue; }
      slow_end(ts0, "tick: cuCtxPushCurrent");
      if (!relaxed) { /* once, with a context current: this thread never takes part in a tenant's capture */
        vgpu_capture_relax();
        relaxed = 1;
      }
      ts0 = slow_t0();
      settle_idle_streams(rt, h);
      slow_end(ts0, "tick: settle_idle_streams");
      if ((epoch % 100) == 1 && rt->lim_h->util_source != VGPU_SRC_NVML) refresh_process_count(rt);
      if ((epoch % 100) == 0 && vgpu_log_level() >= VL_VERBOSE) {
        vgpu_lim_host_t *H = rt->lim_h;
        VLOG(VL_VERBOSE, "limiter host %d: steps %llu user %d (queue %d sm %d) share %lld bucket %lld granted %lld consumed %lld "
             "nproc %d governor %u slot0 launched %llu done %llu", h, (unsigned long long)H->steps, H->user_current,
             H->queue_busy_pct, H->sm_active_pct, (long long)H->share_mirror, (long long)H->bucket_mirror,
             (long long)H->granted_mirror, (long long)H->consumed, H->ext_sys_process_num, H->ctl_state,
             (unsigned long long)H->launched[0], (unsigned long long)H->done[0]);
      }
      if (rt->lim_h->util_source == VGPU_SRC_NVML && !g_governor_mode) {
        if (g_backlog[h].count) backlog_replay(rt, h);
        /* the reference's cadence and the reference's reading: once per control period publish
         * the samples and let one small CTA fold them and refill the bucket */
        if (epoch % g_period_ticks == 0) {
          exchange_with_node_agent(rt, h);
          ts0 = slow_t0();
          g_tick_in_cuda = 0; /* NVML and file reads: nothing the tenant's CUDA calls can hold up */
          unsigned block = publish_utilization(h, rt->u_req);
          g_tick_in_cuda = 1;
          slow_end(ts0, "tick: publish_utilization (NVML)");
          if (vgpu_log_level() >= VL_VERBOSE) {
            const vgpu_util_req_t *U = rt->u_req;
            const vgpu_lim_host_t *H = rt->lim_h;
            int raw = 0;
            for (uint32_t i = 0; U->status == VGPU_UTIL_SAMPLES && i < U->n_samples; i++)
              if (U->samples[i].ts_us >= U->checktime_us) raw += (int)(U->samples[i].sm <= 100 ? U->samples[i].sm : 0);
            struct timespec tn;
            clock_gettime(CLOCK_MONOTONIC, &tn);
            VLOG(VL_VERBOSE, "t=%ld.%03ld host device %d: publication status %u nproc %d samples %u fresh-sm-sum %d | previous step: user util: %d "
                 "sys util: %d share: %lld bucket: %lld up_limit: %d steps: %llu", (long)tn.tv_sec, tn.tv_nsec / 1000000, h, U->status, U->sys_process_num, U->n_samples, raw,
                 H->user_current, H->sys_current, (long long)H->share_mirror, (long long)H->bucket_mirror, H->up_limit_mirror,
                 (unsigned long long)H->steps);
          }
          void *params[] = {&rt->lim_d, &rt->lim_h_d, &rt->u_req_d};
          ts0 = slow_t0();
          CUresult r = VGPU_CAPCHK(R.cuLaunchKernel(rt->k_refill, 1, 1, 1, block, 1, 1, 0, rt->s_stream, params, NULL));
          slow_end(ts0, "tick: refill launch");
          if (r == CUDA_SUCCESS) {
            fails = 0;
            vgpu_metric_add(h, VM_SAMPLER_LAUNCHES, 1);
          } else if (++fails == 25) {
            VLOG(VL_ERROR, "refill launch keeps failing (%d: %s); opening the gate", r, vgpu_cu_err(r));
            rt->memops64 = -1;
            rt->lim_h->granted_mirror = (long long)1 << 60;
          }
        }
      } else if (g_governor_mode) {
        /* the governor owns the queue signal and the controller; the per-SM probe is only
         * needed when the controller is asked to look at SM activity */
        if (rt->lim_h->util_source != VGPU_SRC_QUEUE && rt->lim_h->util_source != VGPU_SRC_NVML &&
            R.cuStreamQuery(rt->p_stream) == CUDA_SUCCESS) {
          uint32_t ep = epoch, never = VGPU_SAMPLER_PROBE_ONLY, none = 0;
          void *params[] = {&rt->lim_d, &rt->lim_h_d, &g_window_us, &g_interval_us, &never, &ep, &none};
          unsigned grid = rt->sm_num > 0 ? (unsigned)rt->sm_num : 148u;
          if (R.cuLaunchKernel(rt->k_sampler, grid, 1, 1, 128, 1, 1, 0, rt->p_stream, params, NULL) == CUDA_SUCCESS)
            vgpu_metric_add(h, VM_SAMPLER_LAUNCHES, 1);
        }
      } else if (!tenant_activity(rt, &g_skipped[h])) {
        /* Nothing of the tenant is executing: a window now could only record "idle".  Skip it -
         * an idle or fully throttled tenant should not keep a kernel of ours on a GPU it shares -
         * and let the next launch account for the skipped ticks (kernels.cu).  While streams are
         * parked one launch per control period is kept, because only the controller releases them. */
        vgpu_metric_add(h, VM_SAMPLER_SKIPPED, 1);
      } else if (VGPU_CAPCHK(R.cuStreamQuery(rt->s_stream)) == CUDA_SUCCESS) {
        uint32_t ep = epoch, skipped = g_skipped[h];
        /* while a tenant thread waits for the device to go idle, keep the sampler's residency
         * negligible so the wait is not stretched by it */
        uint32_t window = g_sync_waiters > 0 ? 200u : g_window_us;
        void *params[] = {&rt->lim_d, &rt->lim_h_d, &window, &g_interval_us, &g_period_ticks, &ep, &skipped};
        /* the queue signal is one warp's work: only the per-SM probe needs a CTA on every SM */
        unsigned grid = rt->lim_h->util_source == VGPU_SRC_QUEUE ? 1u : rt->sm_num > 0 ? (unsigned)rt->sm_num : 148u;
        CUresult r = VGPU_CAPCHK(R.cuLaunchKernel(rt->k_sampler, grid, 1, 1, 128, 1, 1, 0, rt->s_stream, params, NULL));
        if (r == CUDA_SUCCESS) {
          fails = 0;
          g_skipped[h] = 0;
          vgpu_metric_add(h, VM_SAMPLER_LAUNCHES, 1);
        } else if (++fails == 100) {
          /* fail open: never leave tenant streams parked on a bucket nobody refills */
          VLOG(VL_ERROR, "sampler launch keeps failing (%d: %s); opening the gate", r, vgpu_cu_err(r));
          rt->memops64 = -1;
          rt->lim_h->granted_mirror = (long long)1 << 60;
        }
      }
      CUcontext dummy;
      R.cuCtxPopCurrent_v2(&dummy);
      g_tick_in_cuda = 0;
    }
    g_tick_gen++;
    uint64_t after = tick_ns - before;
    struct timespec rest = {(time_t)(after / 1000000000ull), (long)(after % 1000000000ull)};
    nanosleep(&rest, NULL);
  }
}

/* ------------------------------------------------------------------ watchdog
 * Never calls into the driver, so nothing can lock it out.  Condition for a loan: some stream is
 * parked (its newest launch is not admitted), the controller's step counter has not moved, and
 * the tick thread has spent g_watchdog_ms of that time INSIDE the CUDA driver (g_tick_in_cuda) -
 * i.e. it is being kept out by a blocking call of the tenant.  Time the tick thread spends asleep
 * or in NVML does not count: a slow nvmlDeviceGetProcessUtilization (50-110 ms were measured next
 * to a tenant that issues synchronous copies) delays the reference's watcher by the same amount
 * while its throttled thread sleeps in the hook, so lending then would hand out tokens the
 * reference does not.  A tick thread that stopped stepping for any other reason is caught by a
 * backstop of 12 x g_watchdog_ms (fail open).  The loan is the newest parked ticket: everything
 * queued behind the gate (at most GATED_RUNAHEAD launches per stream) is released at once, which
 * is what the blocked driver call is waiting for. */
static uint32_t g_watchdog_ms = 170;
static pthread_t g_wd_tid;
static volatile int g_wd_running;
static void *watchdog_main(void *arg) {
  (void)arg;
  unsigned long long seen_steps[VGPU_MAX_DEVICES] = {0};
  uint32_t stalled_ms[VGPU_MAX_DEVICES] = {0}, parked_ms[VGPU_MAX_DEVICES] = {0};
  const uint32_t step_ms = 10;
  for (;;) {
    struct timespec nap = {0, (long)step_ms * 1000000L};
    nanosleep(&nap, NULL);
    if (g_tick_epoch != vgpu_fork_epoch + 1) return NULL;
    for (int h = 0; h < VGPU_MAX_DEVICES; h++) {
      if (!g_tick_devices[h]) continue;
      vgpu_dev_rt *rt = vgpu_rt_peek(h);
      if (!rt || rt->memops64 < 0) continue;
      vgpu_lim_host_t *H = rt->lim_h;
      long long granted = H->granted_mirror, newest = 0;
      int parked = 0;
      for (uint32_t i = 0; i < VGPU_STREAM_SLOTS; i++) {
        unsigned long long d = H->done[i], l = H->launched[i];
        if (l <= d) continue;
        long long tl = H->ticket[i][l & (VGPU_TICKET_RING - 1)];
        if (granted - tl < 0) {
          if (!parked || tl - newest > 0) newest = tl;
          parked = 1;
        }
      }
      unsigned long long st = H->steps;
      if (!parked || st != seen_steps[h]) {
        seen_steps[h] = st;
        stalled_ms[h] = parked_ms[h] = 0;
        continue;
      }
      parked_ms[h] += step_ms;
      if (g_tick_in_cuda) stalled_ms[h] += step_ms;
      if (stalled_ms[h] < g_watchdog_ms && parked_ms[h] < 12 * g_watchdog_ms) continue;
      const int locked_out = stalled_ms[h] >= g_watchdog_ms;
      stalled_ms[h] = parked_ms[h] = 0;
      H->release_floor = newest;
      __sync_synchronize();
      H->release_pending = 1;
      if (newest - H->granted_mirror > 0) H->granted_mirror = newest;
      vgpu_metric_add(h, VM_WATCHDOG_LOANS, 1);
      VLOG(VL_WARNING, "host device %d: streams are parked and the controller has not stepped: %s; lent tokens up to ticket %lld", h,
           locked_out ? "the refill thread has been waiting inside the CUDA driver (a blocking call of the tenant holds it)"
                      : "the refill thread made no progress for 12 watchdog periods",
           newest);
    }
    g_wd_gen++;
  }
}

/* Process exit: stop launching into a context the CUDA runtime is about to tear down.  Registered
 * with atexit() after the runtime's own handlers, hence run before them. */
static pthread_t g_tick_tid;
static volatile int g_tick_running;
static void tick_stop(void) {
  /* retire every governor: the runtime is about to tear the context down */
  for (int h = 0; h < VGPU_MAX_DEVICES; h++) {
    vgpu_dev_rt *rt = vgpu_rt_peek(h);
    if (!rt || !rt->lim_h) continue;
    __sync_fetch_and_add(&rt->lim_h->quit, 1u);
    struct timespec nap = {0, 100000};
    for (int i = 0; i < 3000 && rt->lim_h->ctl_state != 0; i++) nanosleep(&nap, NULL);
  }
  if (!g_tick_running || g_tick_epoch != vgpu_fork_epoch + 1) return;
  g_tick_epoch = 0; /* tick_main leaves its loop at the next check */
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  ts.tv_sec += 1;
  pthread_timedjoin_np(g_tick_tid, NULL, &ts);
  g_tick_running = 0;
  if (g_wd_running) {
    pthread_timedjoin_np(g_wd_tid, NULL, &ts);
    g_wd_running = 0;
  }
}

static void tick_start(void) {
  g_window_us = env_u32("VGPU_B200_SAMPLER_WINDOW_US", 500);
  g_interval_us = env_u32("VGPU_B200_SAMPLER_INTERVAL_US", 50);
  g_period_ticks = env_u32("VGPU_B200_PERIOD_TICKS", 8);
  g_tick_ms = env_u32("VGPU_B200_TICK_MS", 10);
  if (!g_period_ticks) g_period_ticks = 1;
  if (!g_tick_ms) g_tick_ms = 1;
  g_gov_interval_us = env_u32("VGPU_B200_GOVERNOR_INTERVAL_US", 50);
  g_gov_period_us = env_u32("VGPU_B200_PERIOD_US", 80000);
  g_gov_idle_us = env_u32("VGPU_B200_GOVERNOR_IDLE_US", 5000);
  g_governor_mode = env_u32("VGPU_B200_GOVERNOR", 0) != 0;
  g_watchdog_ms = env_u32("VGPU_B200_WATCHDOG_MS", 170);
  if (g_watchdog_ms < 2 * g_tick_ms * g_period_ticks) g_watchdog_ms = 2 * g_tick_ms * g_period_ticks + 10; /* never sooner than two control periods */
  g_skip_idle = env_u32("VGPU_B200_SKIP_IDLE_WINDOWS", 1) != 0;
  g_tick_epoch = vgpu_fork_epoch + 1;
  if (pthread_create(&g_tick_tid, NULL, tick_main, NULL) == 0) {
    pthread_setname_np(g_tick_tid, "vgpu_b200_tick");
    g_tick_running = 1;
    atexit(tick_stop);
  }
  if (pthread_create(&g_wd_tid, NULL, watchdog_main, NULL) == 0) {
    pthread_setname_np(g_wd_tid, "vgpu_b200_wdog");
    g_wd_running = 1;
  }
}

/* Same start-up strictness as the reference's initialization() / init_device_cuda_cores
 * (cuda_hook.c:487-577): every CUDA-visible device must be described by the config and be
 * known to NVML, otherwise the process is terminated. */
static pthread_once_t g_verify_once = PTHREAD_ONCE_INIT;
static void verify_devices(void) {
  if (!R.cuInit || R.cuInit(0) != CUDA_SUCCESS) {
    VLOG(VL_ERROR, "initialization of sm watcher failed");
    return;
  }
  int n = 0;
  CUresult r = R.cuDeviceGetCount ? R.cuDeviceGetCount(&n) : CUDA_ERROR_NOT_FOUND;
  if (r) VLOG(VL_FATAL, "cuDeviceGetCount call failed, return %d, str: %s", r, vgpu_cu_err(r));
  for (int i = 0; i < n; i++) {
    CUdevice dev;
    r = R.cuDeviceGet(&dev, i);
    if (r) VLOG(VL_FATAL, "cuDeviceGet call failed, cuda device %d, return %d, str %s", i, r, vgpu_cu_err(r));
    if (vgpu_host_index_of_cuda(dev) < 0)
      VLOG(VL_FATAL, "cuda device %d cannot find the corresponding host device", dev);
    if (vgpu_nvml_index_of_cuda(dev) < 0)
      VLOG(VL_FATAL, "cuda device %d cannot find the corresponding nvml device", dev);
  }
}

static void tick_start(void);
void vgpu_limiter_start(void) {
  pthread_once(&g_verify_once, verify_devices);
  /* reference: initialization() spawns watch_util_bt_N threads at the first successful cuInit
   * (cuda_hook.c:566-577).  Here the thread is created lazily by the first limited launch,
   * because the sampler needs the tenant's context; this entry point only re-arms after fork
   * (the reference does not - SURVEY.md Appendix B.13). */
  if (g_tick_epoch && g_tick_epoch != vgpu_fork_epoch + 1) {
    memset((void *)g_tick_devices, 0, sizeof g_tick_devices);
    memset((void *)g_slots, 0, sizeof g_slots);
    g_tick_once = (pthread_once_t)PTHREAD_ONCE_INIT;
    g_tick_epoch = 0;
    memset(g_backlog, 0, sizeof g_backlog);
  }
  /* like the reference's initialization(): the watcher runs from the first successful cuInit */
  int n = 0, any = 0;
  if (R.cuDeviceGetCount && R.cuDeviceGetCount(&n) == CUDA_SUCCESS)
    for (int i = 0; i < n; i++) {
      CUdevice dev;
      if (R.cuDeviceGet(&dev, i) != CUDA_SUCCESS) continue;
      int h = vgpu_host_index_of_cuda(dev);
      if (h >= 0 && G_cfg->devices[h].core_limit) {
        g_tick_devices[h] = 1;
        any = 1;
      }
    }
  if (any) pthread_once(&g_tick_once, tick_start);
}

/* ------------------------------------------------------------------ governor */

/* Called by the launch hook after it has published its launch.  Dekker pair with the governor's
 * retire protocol (kernels.cu): publish, fence, then look at the state. */
static inline void governor_ensure(vgpu_dev_rt *rt, int h) {
  vgpu_lim_host_t *H = rt->lim_h;
  __sync_synchronize();
  uint32_t st = H->ctl_state;
  if (likely(st == 1)) return;
  for (int spins = 0; st == 2 && spins < 200000; spins++) { /* leaving: its verdict takes microseconds */
    __builtin_ia32_pause();
    st = H->ctl_state;
  }
  if (st == 1) return;
  if (!__sync_bool_compare_and_swap(&H->ctl_state, st, 1u)) return; /* another thread is starting it */
  CUcontext cur = NULL;
  int pushed = 0;
  if (R.cuCtxGetCurrent(&cur) == CUDA_SUCCESS && cur != rt->ctx && R.cuCtxPushCurrent_v2(rt->ctx) == CUDA_SUCCESS) pushed = 1;
  void *params[] = {&rt->lim_d, &rt->lim_h_d, &g_gov_interval_us, &g_gov_period_us, &g_gov_idle_us};
  CUresult r = R.cuLaunchKernel(rt->k_governor, 1, 1, 1, 32, 1, 1, 0, rt->s_stream, params, NULL);
  if (pushed) R.cuCtxPopCurrent_v2(&cur);
  if (likely(r == CUDA_SUCCESS)) {
    vgpu_metric_add(h, VM_SAMPLER_LAUNCHES, 1);
  } else {
    H->ctl_state = 0;
    VLOG(VL_ERROR, "governor launch failed (%d: %s); launches stay un-gated until it succeeds", r, vgpu_cu_err(r));
    rt->memops64 = -1; /* no gate without a governor */
  }
}

/* Bracket around a driver call that synchronises the whole device (it would otherwise wait for
 * the resident governor).  begin: give every launch train its completion marker so the queue
 * state the governor sees is exact, then announce the synchronise; the governor leaves as soon as
 * nothing is parked.  end: withdraw the announcement and bring the governor back if tenant work
 * is (or was, when it left) still executing, so that the time it was away is attributed. */
static __thread unsigned long long t_sync_snap[VGPU_STREAM_SLOTS];
static inline void enqueue_marker(vgpu_dev_rt *rt, int h, uint32_t slot, unsigned long long seq, CUstream s, int ptsz);

/* The tenant is tearing a context down: make sure neither background thread is inside (or will
 * enter) this device's runtime.  Both re-read g_tick_devices[] at the top of every iteration, so
 * one completed iteration after the flag is cleared is a quiescence point. */
void vgpu_limiter_detach(int h) {
  if (h < 0 || h >= VGPU_MAX_DEVICES || !g_tick_devices[h]) return;
  g_tick_devices[h] = 0;
  __sync_synchronize();
  if (g_tick_epoch != vgpu_fork_epoch + 1) return;
  unsigned long t0 = g_tick_gen, w0 = g_wd_gen;
  struct timespec nap = {0, 1000000};
  for (int i = 0; i < 2000; i++) { /* 2 s: a tick thread stuck in the driver must not wedge the tenant */
    if ((!g_tick_running || g_tick_gen != t0) && (!g_wd_running || g_wd_gen != w0)) break;
    nanosleep(&nap, NULL);
  }
}

void vgpu_limiter_attach(int h, int forget_streams) {
  if (h < 0 || h >= VGPU_MAX_DEVICES) return;
  if (forget_streams) {
    memset((void *)g_slots[h], 0, sizeof g_slots[h]);
    return; /* the next limited launch registers the device again */
  }
  g_tick_devices[h] = 1;
}

void vgpu_limiter_quiesce(vgpu_dev_rt *rt) {
  if (!rt || !rt->lim_h) return;
  vgpu_lim_host_t *H = rt->lim_h;
  int h = rt->host_index;
  if (h >= 0 && g_tick_devices[h]) {
    for (uint32_t i = 0; i < VGPU_STREAM_SLOTS; i++) {
      slot_t *sl = &g_slots[h][i];
      unsigned long long l = H->launched[i];
      t_sync_snap[i] = l;
      if (sl->key <= SLOT_TOMB || l <= H->done[i] || sl->marked >= l || rt->memops64 <= 0) continue;
      if (i == VGPU_STREAM_SLOTS - 1 || (sl->ptsz && !sl->stream)) continue; /* marked per launch already */
      enqueue_marker(rt, h, i, l, sl->stream, sl->ptsz);
    }
  }
  __sync_fetch_and_add(&H->quit, 1u);
}

void vgpu_limiter_resume(vgpu_dev_rt *rt, int everything_completed) {
  if (!rt || !rt->lim_h) return;
  vgpu_lim_host_t *H = rt->lim_h;
  int h = rt->host_index;
  __sync_fetch_and_sub(&H->quit, 1u);
  if (h < 0 || !g_tick_devices[h]) return;
  int outstanding = 0;
  for (uint32_t i = 0; i < VGPU_STREAM_SLOTS; i++) {
    if (everything_completed) { /* whatever had been launched before the call has finished */
      unsigned long long d;
      while ((d = H->done[i]) < t_sync_snap[i] && !__sync_bool_compare_and_swap(&H->done[i], d, t_sync_snap[i])) {}
    }
    if (H->launched[i] > H->done[i]) outstanding = 1;
  }
  if (g_governor_mode && (outstanding || H->gov_left_busy)) governor_ensure(rt, h);
}

/* ------------------------------------------------------------------ admission */
#define GATED_RUNAHEAD 96u /* launches (x3 stream ops) allowed to queue behind a closed gate */
typedef struct {
  vgpu_dev_rt *rt;
  uint32_t slot;
  unsigned long long seq;
  int ptsz;
  int capturing; /* the launch was recorded into a graph (found out at the gate) */
} admit_t;

static inline void enqueue_marker(vgpu_dev_rt *rt, int h, uint32_t slot, unsigned long long seq, CUstream s, int ptsz) {
  CUdeviceptr addr = rt->lim_h_d + offsetof(vgpu_lim_host_t, done) + (CUdeviceptr)slot * sizeof(unsigned long long);
  CUresult (*wr)(CUstream, CUdeviceptr, cuuint64_t, unsigned) =
      (ptsz && R.cuStreamWriteValue64_v2_ptsz) ? R.cuStreamWriteValue64_v2_ptsz : R.cuStreamWriteValue64_v2;
  if (likely(VGPU_CAPCHK(wr(s, addr, (cuuint64_t)seq, 0)) == CUDA_SUCCESS)) g_slots[h][slot].marked = seq;
  else rt->memops64 = 0;
}

/* Is `s` recording into a graph right now?  Only asked where the library is about to enqueue a
 * stream operation of its own (marker, gate) - those must not be recorded, they would replay
 * stale values - never on the per-launch fast path.  The legacy NULL stream cannot capture. */
static inline int stream_capturing(CUstream s, int ptsz) {
  int capturing = 0;
  if ((s != NULL || ptsz) && R.cuStreamIsCapturing) VGPU_CAPCHK(R.cuStreamIsCapturing(s, &capturing));
  return capturing;
}

/* returns 0 when the launch should simply be forwarded */
static inline int admit(admit_t *a, unsigned gx, unsigned gy, unsigned gz, CUstream s, int ptsz) {
  CUdevice dev;
  if (unlikely(!G_cfg)) vgpu_boot(); /* entry reached without cuInit/dlsym going through us */
  if (unlikely(g_tick_epoch && g_tick_epoch != vgpu_fork_epoch + 1)) vgpu_limiter_start(); /* forked child */
  if (unlikely(R.cuCtxGetDevice(&dev) != CUDA_SUCCESS)) return -1;
  int h = vgpu_host_index_of_cuda(dev);
  if (h < 0 || !G_cfg->devices[h].core_limit) return 0;
  vgpu_dev_rt *rt = vgpu_rt_get(h, dev);
  if (unlikely(!rt)) { /* bring-up failed: logged there, retried after its back-off */
    /* There is no CPU enforcement path to fall back to.  Default: the launch passes un-throttled until the retry
     * (availability first, like a reference whose watcher thread failed to start).  VGPU_B200_FAIL_CLOSED=1 (a †
     * tunable: control plane's b200.tunables, or the env of an env-configured tenant): refuse it instead. */
    static int fail_closed = -1;
    if (fail_closed < 0) fail_closed = env_u32("VGPU_B200_FAIL_CLOSED", 0) != 0;
    return fail_closed ? -2 : 0;
  }
  if (unlikely(!g_tick_devices[h])) {
    g_tick_devices[h] = 1;
    pthread_once(&g_tick_once, tick_start);
  }
  vgpu_lim_host_t *H = rt->lim_h;
  /* the reference multiplies the three unsigned dims in 32 bits and passes the result as int */
  long long cost = (long long)(int)(gx * gy * gz);
  a->rt = rt;
  a->ptsz = ptsz;
  a->slot = slot_of(h, s, ptsz);
  slot_t *sl = &g_slots[h][a->slot];
  /* bounded run-ahead: the ticket ring holds VGPU_TICKET_RING outstanding launches per stream
   * (signed difference: several threads may share a slot - the legacy stream, the per-thread
   * default streams, the overflow slot - and `done` may be a moment ahead of a stale read) */
  if (unlikely((long long)(H->launched[a->slot] + 1 - H->done[a->slot]) >= (long long)VGPU_TICKET_RING - 64)) {
    /* wait for the stream to drain a little; never forever - if completion markers stopped
     * arriving (driver refused the mem-op, context torn down) fall back to marker-less mode */
    /* a marker is only added here when none is in flight (the regular one-per-MARK_EVERY markers
     * normally guarantee progress; every extra one is another operation for the GPU front-end) */
    if (sl->marked <= H->done[a->slot] && sl->marked < H->launched[a->slot] && rt->memops64 > 0 && !stream_capturing(s, ptsz))
      enqueue_marker(rt, h, a->slot, H->launched[a->slot], s, ptsz);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    while ((long long)(H->launched[a->slot] + 1 - H->done[a->slot]) >= (long long)VGPU_TICKET_RING - 64) {
      sched_yield();
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if (t1.tv_sec - t0.tv_sec >= 5) {
        VLOG(VL_ERROR, "completion markers stalled on host device %d; disabling stream mem-ops", h);
        rt->memops64 = 0;
        H->done[a->slot] = H->launched[a->slot];
        break;
      }
    }
  }
  /* ticket, sequence number and ring entry of one launch are taken together, so tickets grow
   * monotonically inside a slot even when several threads launch into it */
  slot_lock(sl);
  long long ticket = __sync_fetch_and_add(&H->consumed, cost);
  unsigned long long seq = H->launched[a->slot] + 1;
  H->ticket[a->slot][seq & (VGPU_TICKET_RING - 1)] = ticket;
  __atomic_store_n(&H->launched[a->slot], seq, __ATOMIC_RELEASE); /* ticket first, then the sequence */
  slot_unlock(sl);
  a->seq = seq;
  if (unlikely(g_governor_mode)) governor_ensure(rt, h);
  if (H->granted_mirror - ticket < 0 && rt->memops64 >= 0) {
    if (stream_capturing(s, ptsz)) { /* graph capture: tokens are paid at capture time, like the
                                        reference; no gate or marker nodes are recorded */
      a->capturing = 1;
      return 1;
    }
    /* bucket empty: park the *stream* on the bucket word (its host-visible mirror, so that the
     * watchdog can lend tokens without the driver), not the CPU thread */
    vgpu_metric_add(h, VM_RATE_GATED, 1);
    /* Bounded run-ahead.  A parked stream must never be allowed to fill the driver's hardware
     * queue: a launch call that blocks inside the driver for queue space holds the context lock,
     * and the tick thread could then no longer launch the refill whose controller is the only
     * thing that can release the stream.  So once GATED_RUNAHEAD launches are queued behind the
     * gate, wait here - in user space, in ~20 us steps - for the stream to drain or for tokens.
     * (The reference blocks the thread for every throttled launch, in 10 ms steps.) */
    if (unlikely((long long)(seq - H->done[a->slot]) > (long long)GATED_RUNAHEAD)) {
      struct timespec nap = {0, 20000};
      for (int spins = 0; (long long)(seq - H->done[a->slot]) > (long long)GATED_RUNAHEAD && H->granted_mirror - ticket < 0; spins++) {
        if (spins < 64) sched_yield();
        else nanosleep(&nap, NULL);
        if (unlikely(rt->memops64 <= 0)) break;
      }
    }
    /* make the controller's view exact at the gate: everything before this launch gets its
     * marker now, so the oldest unfinished launch it will see is this (parked) one */
    uint64_t tm0 = slow_t0();
    if (likely(rt->memops64 > 0) && sl->marked < seq - 1) enqueue_marker(rt, h, a->slot, seq - 1, s, ptsz);
    slow_end(tm0, "launch: marker before the gate");
    if (likely(rt->memops64 > 0)) {
      CUresult (*wait)(CUstream, CUdeviceptr, cuuint64_t, unsigned) =
          (ptsz && R.cuStreamWaitValue64_v2_ptsz) ? R.cuStreamWaitValue64_v2_ptsz : R.cuStreamWaitValue64_v2;
      uint64_t ts0 = slow_t0();
      CUresult wr = VGPU_CAPCHK(wait(s, rt->lim_h_d + offsetof(vgpu_lim_host_t, granted_mirror), (cuuint64_t)ticket, VCU_WAIT_GEQ));
      slow_end(ts0, "launch: cuStreamWaitValue64 (gate)");
      if (unlikely(wr != CUDA_SUCCESS)) {
        VLOG(VL_ERROR, "cuStreamWaitValue64 failed (%d: %s); falling back to the gate kernel", wr, vgpu_cu_err(wr));
        rt->memops64 = 0;
      }
    }
    if (unlikely(rt->memops64 == 0)) {
      CUdeviceptr gp = rt->lim_h_d + offsetof(vgpu_lim_host_t, granted_mirror);
      uint32_t timeout_ms = 2000;
      void *params[] = {&gp, &ticket, &timeout_ms};
      (ptsz && R.cuLaunchKernel_ptsz ? R.cuLaunchKernel_ptsz : R.cuLaunchKernel)(
          rt->k_gate, 1, 1, 1, 1, 1, 1, 0, s, params, NULL);
    }
  } else {
    vgpu_metric_add(h, VM_RATE_FAST, 1);
  }
  return 1;
}

static inline void mark_done(const admit_t *a, CUstream s) {
  vgpu_dev_rt *rt = a->rt;
  if (!rt) return;
  int h = rt->host_index;
  slot_t *sl = &g_slots[h][a->slot];
  if (likely(rt->memops64 > 0)) {
    /* a completion marker costs a driver call (~3 us): dense trains share one per MARK_EVERY
     * launches; the overflow slot and other threads' default streams get one per launch (the
     * tick thread cannot settle those); the on-device queue signal also wants one behind every
     * isolated launch.  With the default NVML reading the markers only bound the run-ahead and
     * locate the first parked launch, so the dense rule is enough. */
    int due = a->seq - sl->marked >= MARK_EVERY || a->slot == VGPU_STREAM_SLOTS - 1 || (sl->ptsz && !sl->stream);
    if (rt->lim_h->util_source != VGPU_SRC_NVML) {
      unsigned long long now = rdtsc();
      due |= (now - sl->last_tsc) > SPARSE_TSC;
      sl->last_tsc = now;
    }
    if (likely(!due && !a->capturing)) return;
    if (a->capturing || stream_capturing(s, a->ptsz)) {
      /* recorded into a graph, not executed: nothing is in flight for this sequence number */
      unsigned long long d;
      while ((d = rt->lim_h->done[a->slot]) < a->seq && !__sync_bool_compare_and_swap(&rt->lim_h->done[a->slot], d, a->seq)) {}
      return;
    }
    uint64_t tm0 = slow_t0();
    enqueue_marker(rt, h, a->slot, a->seq, s, a->ptsz);
    slow_end(tm0, "launch: completion marker");
    if (likely(rt->memops64 > 0)) return;
  }
  /* no completion signal available: treat as instantaneous (monotonic: slots can be shared) */
  unsigned long long d;
  while ((d = rt->lim_h->done[a->slot]) < a->seq && !__sync_bool_compare_and_swap(&rt->lim_h->done[a->slot], d, a->seq)) {}
}

#define LIMITED_LAUNCH(gx, gy, gz, stream, ptsz, CALL)          \
  do {                                                          \
    admit_t a_ = {0};                                           \
    int st_ = admit(&a_, (gx), (gy), (gz), (stream), (ptsz));   \
    if (st_ < 0) return st_ == -2 ? CUDA_ERROR_NOT_SUPPORTED : CUDA_ERROR_INVALID_CONTEXT; \
    uint64_t ts_ = slow_t0();                                   \
    CUresult r_ = (CALL);                                       \
    slow_end(ts_, "launch: driver call");                       \
    if (st_ > 0) mark_done(&a_, (stream));                      \
    return r_;                                                  \
  } while (0)

VGPU_EXPORT CUresult cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx,
                                    unsigned by, unsigned bz, unsigned smem, CUstream s, void **params,
                                    void **extra) {
  if (unlikely(!R.cuLaunchKernel)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH(gx, gy, gz, s, 0, R.cuLaunchKernel(f, gx, gy, gz, bx, by, bz, smem, s, params, extra));
}

VGPU_EXPORT CUresult cuLaunchKernel_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx,
                                         unsigned by, unsigned bz, unsigned smem, CUstream s,
                                         void **params, void **extra) {
  if (unlikely(!R.cuLaunchKernel_ptsz)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH(gx, gy, gz, s, 1, R.cuLaunchKernel_ptsz(f, gx, gy, gz, bx, by, bz, smem, s, params, extra));
}

VGPU_EXPORT CUresult cuLaunchKernelEx(const vcu_launch_config_t *c, CUfunction f, void **params, void **extra) {
  if (unlikely(!R.cuLaunchKernelEx)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH(c->gridDimX, c->gridDimY, c->gridDimZ, c->hStream, 0, R.cuLaunchKernelEx(c, f, params, extra));
}

VGPU_EXPORT CUresult cuLaunchKernelEx_ptsz(const vcu_launch_config_t *c, CUfunction f, void **params,
                                           void **extra) {
  if (unlikely(!R.cuLaunchKernelEx_ptsz)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH(c->gridDimX, c->gridDimY, c->gridDimZ, c->hStream, 1,
                 R.cuLaunchKernelEx_ptsz(c, f, params, extra));
}

VGPU_EXPORT CUresult cuLaunchCooperativeKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz,
                                               unsigned bx, unsigned by, unsigned bz, unsigned smem,
                                               CUstream s, void **params) {
  if (unlikely(!R.cuLaunchCooperativeKernel)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH(gx, gy, gz, s, 0, R.cuLaunchCooperativeKernel(f, gx, gy, gz, bx, by, bz, smem, s, params));
}

VGPU_EXPORT CUresult cuLaunchCooperativeKernel_ptsz(CUfunction f, unsigned gx, unsigned gy, unsigned gz,
                                                    unsigned bx, unsigned by, unsigned bz, unsigned smem,
                                                    CUstream s, void **params) {
  if (unlikely(!R.cuLaunchCooperativeKernel_ptsz)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH(gx, gy, gz, s, 1,
                 R.cuLaunchCooperativeKernel_ptsz(f, gx, gy, gz, bx, by, bz, smem, s, params));
}

/* legacy launch API: the grid comes from the call, the block shape from cuFuncSetBlockShape
 * (cuda_hook.c:1883-2002); only the grid size is charged */
VGPU_EXPORT CUresult cuLaunch(CUfunction f) {
  if (unlikely(!R.cuLaunch)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH(1, 1, 1, NULL, 0, R.cuLaunch(f));
}
VGPU_EXPORT CUresult cuLaunchGrid(CUfunction f, int w, int h) {
  if (unlikely(!R.cuLaunchGrid)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH((unsigned)(w * h), 1, 1, NULL, 0, R.cuLaunchGrid(f, w, h));
}
VGPU_EXPORT CUresult cuLaunchGridAsync(CUfunction f, int w, int h, CUstream s) {
  if (unlikely(!R.cuLaunchGridAsync)) return CUDA_ERROR_NOT_FOUND;
  LIMITED_LAUNCH((unsigned)(w * h), 1, 1, s, 0, R.cuLaunchGridAsync(f, w, h, s));
}
VGPU_EXPORT CUresult cuFuncSetBlockShape(CUfunction f, int x, int y, int z) {
  CUdevice dev; /* the reference only caches the shape for logging; the driver call needs a ctx */
  if (!R.cuCtxGetDevice || R.cuCtxGetDevice(&dev) != CUDA_SUCCESS) return CUDA_ERROR_INVALID_CONTEXT;
  return R.cuFuncSetBlockShape ? R.cuFuncSetBlockShape(f, x, y, z) : CUDA_ERROR_NOT_FOUND;
}

/* ------------------------------------------------------------------ blocking calls of a throttled tenant
 * In the reference a throttled thread sleeps inside the launch hook (cuda_hook.c:322-326), so it
 * never reaches a blocking driver call with work still waiting for tokens.  Here the launch
 * returns at once and the *stream* waits, which means the thread can walk into cuCtxSynchronize,
 * cuStreamSynchronize, cuEventSynchronize, a synchronous cuMemcpyDtoH or cuMemFree while its
 * stream is parked - and several of those keep other threads of the process out of the driver
 * for as long as they block, including the tick thread whose refill launch is the only thing
 * that releases the stream (measured: 170 ms stalls ended by watchdog loans, and tenants that hit
 * them more often than their neighbours fell behind by up to 7x).  So the hooked blocking calls
 * first wait in user space - exactly where the reference's thread would be sleeping - until
 * nothing they depend on is parked any more, and only then enter the driver. */
static int slot_parked(const vgpu_lim_host_t *H, uint32_t i) {
  unsigned long long l = H->launched[i], d = H->done[i];
  if (l <= d) return 0;
  return H->granted_mirror - H->ticket[i][l & (VGPU_TICKET_RING - 1)] < 0; /* newest launch not admitted yet */
}

static void wait_until_unparked(vgpu_dev_rt *rt, int h, CUstream s, int ptsz, int everything) {
  if (!rt || !rt->lim_h || h < 0 || !g_tick_devices[h] || rt->memops64 < 0) return;
  const vgpu_lim_host_t *H = rt->lim_h;
  uint32_t only = VGPU_STREAM_SLOTS;
  if (!everything && (s != NULL || ptsz)) { /* the legacy NULL stream synchronises with every blocking stream */
    uintptr_t key = slot_key(s, ptsz);
    for (uint32_t i = 0; i < VGPU_STREAM_SLOTS - 1; i++)
      if (g_slots[h][i].key == key) { only = i; break; }
    if (only == VGPU_STREAM_SLOTS) return; /* never launched into through us: nothing of ours is parked there */
  }
  struct timespec t0, now, nap = {0, 50000};
  clock_gettime(CLOCK_MONOTONIC, &t0);
  uint64_t ts0 = slow_t0();
  for (int spins = 0;; spins++) {
    int parked = 0;
    if (only < VGPU_STREAM_SLOTS) parked = slot_parked(H, only) || slot_parked(H, VGPU_STREAM_SLOTS - 1);
    else
      for (uint32_t i = 0; i < VGPU_STREAM_SLOTS && !parked; i++) parked = slot_parked(H, i);
    if (!parked) { slow_end(ts0, "blocking call: user-space wait for tokens"); return; }
    if (spins < 32) sched_yield();
    else nanosleep(&nap, NULL);
    if ((spins & 255) == 255) {
      clock_gettime(CLOCK_MONOTONIC, &now);
      if (now.tv_sec - t0.tv_sec >= 5) return; /* never for ever: the watchdog covers a dead refill path */
    }
  }
}

static vgpu_dev_rt *current_rt(int *h_out) {
  CUdevice dev;
  *h_out = -1;
  if (unlikely(!G_cfg)) vgpu_boot();
  if (!R.cuCtxGetDevice || R.cuCtxGetDevice(&dev) != CUDA_SUCCESS) return NULL;
  *h_out = vgpu_host_index_of_cuda(dev);
  return vgpu_rt_peek(*h_out);
}

void vgpu_limiter_before_blocking_call(vgpu_dev_rt *rt) { /* cuMemFree and friends (memgate.c) */
  if (rt) wait_until_unparked(rt, rt->host_index, NULL, 0, 1);
}

VGPU_EXPORT CUresult cuStreamSynchronize(CUstream s) {
  int h;
  vgpu_dev_rt *rt = current_rt(&h);
  if (unlikely(!R.cuStreamSynchronize)) return CUDA_ERROR_NOT_FOUND;
  wait_until_unparked(rt, h, s, 0, 0);
  return R.cuStreamSynchronize(s);
}
VGPU_EXPORT CUresult cuStreamSynchronize_ptsz(CUstream s) {
  int h;
  vgpu_dev_rt *rt = current_rt(&h);
  if (unlikely(!R.cuStreamSynchronize_ptsz)) return CUDA_ERROR_NOT_FOUND;
  wait_until_unparked(rt, h, s, 1, 0);
  return R.cuStreamSynchronize_ptsz(s);
}
VGPU_EXPORT CUresult cuEventSynchronize(CUevent e) {
  int h;
  vgpu_dev_rt *rt = current_rt(&h);
  if (unlikely(!R.cuEventSynchronize)) return CUDA_ERROR_NOT_FOUND;
  wait_until_unparked(rt, h, NULL, 0, 1); /* the event may sit behind any stream */
  uint64_t ts0 = slow_t0();
  CUresult r = R.cuEventSynchronize(e);
  slow_end(ts0, "cuEventSynchronize: driver call");
  return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoH_v2(void *dst, CUdeviceptr src, size_t n) {
  int h;
  vgpu_dev_rt *rt = current_rt(&h);
  if (unlikely(!R.cuMemcpyDtoH_v2)) return CUDA_ERROR_NOT_FOUND;
  wait_until_unparked(rt, h, NULL, 0, 1);
  uint64_t ts0 = slow_t0();
  CUresult r = R.cuMemcpyDtoH_v2(dst, src, n);
  slow_end(ts0, "cuMemcpyDtoH: driver call");
  return r;
}
VGPU_EXPORT CUresult cuMemcpyDtoH_v2_ptds(void *dst, CUdeviceptr src, size_t n) {
  int h;
  vgpu_dev_rt *rt = current_rt(&h);
  if (unlikely(!R.cuMemcpyDtoH_v2_ptds)) return CUDA_ERROR_NOT_FOUND;
  wait_until_unparked(rt, h, NULL, 1, 0);
  return R.cuMemcpyDtoH_v2_ptds(dst, src, n);
}

/* the synchronous copies that involve host memory wait for everything on the legacy stream (per-thread stream
 * for _ptds) inside the driver; device-to-device copies are only enqueued and stay forwards */
#define BLOCKING_COPY(name, ptsz, T1, T2)                                  \
  VGPU_EXPORT CUresult name(T1 dst, T2 src, size_t n) {                    \
    int h;                                                                 \
    vgpu_dev_rt *rt = current_rt(&h);                                      \
    if (unlikely(!R.name)) return CUDA_ERROR_NOT_FOUND;                    \
    wait_until_unparked(rt, h, NULL, ptsz, !(ptsz));                       \
    uint64_t ts0 = slow_t0();                                              \
    CUresult r = R.name(dst, src, n);                                      \
    slow_end(ts0, #name ": driver call");                                  \
    return r;                                                              \
  }
BLOCKING_COPY(cuMemcpyHtoD_v2, 0, CUdeviceptr, const void *)
BLOCKING_COPY(cuMemcpyHtoD_v2_ptds, 1, CUdeviceptr, const void *)
BLOCKING_COPY(cuMemcpy, 0, CUdeviceptr, CUdeviceptr)
BLOCKING_COPY(cuMemcpy_ptds, 1, CUdeviceptr, CUdeviceptr)

/* B200 addition: a device-wide synchronise must not wait for the resident governor, and must not
 * block inside the driver while tenant work is still waiting for tokens */
VGPU_EXPORT CUresult cuCtxSynchronize(void) {
  vgpu_boot();
  if (unlikely(!R.cuCtxSynchronize)) return CUDA_ERROR_NOT_FOUND;
  CUdevice dev;
  vgpu_dev_rt *rt = NULL;
  if (R.cuCtxGetDevice && R.cuCtxGetDevice(&dev) == CUDA_SUCCESS) rt = vgpu_rt_peek(vgpu_host_index_of_cuda(dev));
  if (rt) wait_until_unparked(rt, rt->host_index, NULL, 0, 1);
  vgpu_limiter_quiesce(rt);
  __sync_fetch_and_add(&g_sync_waiters, 1);
  uint64_t ts0 = slow_t0();
  CUresult r = R.cuCtxSynchronize();
  slow_end(ts0, "cuCtxSynchronize: driver call");
  __sync_fetch_and_sub(&g_sync_waiters, 1);
  vgpu_limiter_resume(rt, r == CUDA_SUCCESS);
  return r;
}

/* B200 addition: forget a stream's slot when the application destroys it */
VGPU_EXPORT CUresult cuStreamDestroy_v2(CUstream s) {
  vgpu_boot();
  for (int h = 0; h < VGPU_MAX_DEVICES; h++)
    for (uint32_t i = 0; i < VGPU_STREAM_SLOTS - 1; i++) {
      slot_t *sl = &g_slots[h][i];
      if (sl->key > SLOT_TOMB && sl->stream == s) {
        vgpu_dev_rt *rt = vgpu_rt_peek(h);
        slot_lock(sl);
        if (rt) rt->lim_h->done[i] = rt->lim_h->launched[i];
        sl->stream = NULL;
        sl->marked = sl->seen_launched = 0;
        __sync_synchronize();
        sl->key = SLOT_TOMB; /* not 0: streams that probed past this slot keep finding theirs */
        slot_unlock(sl);
      }
    }
  return R.cuStreamDestroy_v2 ? R.cuStreamDestroy_v2(s) : CUDA_ERROR_NOT_FOUND;
}

/* ------------------------------------------------------------------ CUDA graphs (opt-in)
 * The reference meters kernels at capture time only (the launch hooks run while the stream
 * captures) and forwards cuGraphLaunch untouched, so every replay of a graph is free and
 * invisible to its utilisation reading.  With VGPU_B200_GRAPH_LIMIT=1 a replay costs what its
 * kernel nodes would cost as individual launches (sum of gridX*gridY*gridZ, child graphs
 * included), is gated like a launch and carries a completion marker.  Off by default: it changes
 * behaviour relative to the reference. */
#define GRAPH_TAB 2048u
typedef struct { volatile uintptr_t exec; long long cost; } graph_ent;
static graph_ent g_graphs[GRAPH_TAB];
static pthread_mutex_t g_graph_mu = PTHREAD_MUTEX_INITIALIZER;

int vgpu_graph_limit_enabled(void) {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("VGPU_B200_GRAPH_LIMIT");
    on = (e && (*e == '1' || *e == 't' || *e == 'T')) ? 1 : 0;
  }
  return on;
}

static long long graph_cost(CUgraph g, int depth) {
  size_t n = 0;
  if (!R.cuGraphGetNodes || !R.cuGraphNodeGetType || depth > 8) return 0;
  if (R.cuGraphGetNodes(g, NULL, &n) != CUDA_SUCCESS || n == 0) return 0;
  CUgraphNode *nodes = (CUgraphNode *)malloc(n * sizeof *nodes);
  if (!nodes) return 0;
  long long cost = 0;
  if (R.cuGraphGetNodes(g, nodes, &n) == CUDA_SUCCESS) {
    for (size_t i = 0; i < n; i++) {
      int type = -1;
      if (R.cuGraphNodeGetType(nodes[i], &type) != CUDA_SUCCESS) continue;
      if (type == VCU_GRAPH_NODE_KERNEL) {
        /* CUDA_KERNEL_NODE_PARAMS v1 (56 B) and v2 (72 B) both start {CUfunction func; unsigned gridDimX,Y,Z; ...} */
        union { unsigned u[32]; void *align; } p;
        memset(&p, 0, sizeof p);
        CUresult r = R.cuGraphKernelNodeGetParams_v2 ? R.cuGraphKernelNodeGetParams_v2(nodes[i], &p) : CUDA_ERROR_NOT_FOUND;
        if (r != CUDA_SUCCESS && R.cuGraphKernelNodeGetParams) r = R.cuGraphKernelNodeGetParams(nodes[i], &p);
        if (r == CUDA_SUCCESS) cost += (long long)(int)(p.u[2] * p.u[3] * p.u[4]); /* same 32-bit product as a launch */
      } else if (type == VCU_GRAPH_NODE_GRAPH && R.cuGraphChildGraphNodeGetGraph) {
        CUgraph child = NULL;
        if (R.cuGraphChildGraphNodeGetGraph(nodes[i], &child) == CUDA_SUCCESS && child) cost += graph_cost(child, depth + 1);
      }
    }
  }
  free(nodes);
  return cost;
}

static void graph_remember(CUgraphExec exec, long long cost) {
  uintptr_t key = (uintptr_t)exec;
  uint32_t h = (uint32_t)((key >> 4) * 0x9E3779B97F4A7C15ull >> 53) % GRAPH_TAB;
  pthread_mutex_lock(&g_graph_mu);
  for (uint32_t i = 0; i < GRAPH_TAB; i++) {
    graph_ent *e = &g_graphs[(h + i) % GRAPH_TAB];
    if (e->exec == 0 || e->exec == 1 || e->exec == key) { /* empty, tombstone or re-used handle */
      e->cost = cost;
      __sync_synchronize();
      e->exec = key;
      break;
    }
  }
  pthread_mutex_unlock(&g_graph_mu);
}

static int graph_lookup(CUgraphExec exec, long long *cost, int forget) {
  uintptr_t key = (uintptr_t)exec;
  uint32_t h = (uint32_t)((key >> 4) * 0x9E3779B97F4A7C15ull >> 53) % GRAPH_TAB;
  for (uint32_t i = 0; i < GRAPH_TAB; i++) {
    graph_ent *e = &g_graphs[(h + i) % GRAPH_TAB];
    uintptr_t k = e->exec;
    if (k == 0) return 0;
    if (k == key) {
      if (cost) *cost = e->cost;
      if (forget) e->exec = 1;
      return 1;
    }
  }
  return 0;
}

static void graph_instantiated(CUgraphExec *out, CUgraph g, CUresult r) {
  if (r != CUDA_SUCCESS || !out || !*out || !vgpu_graph_limit_enabled()) return;
  long long cost = graph_cost(g, 0);
  graph_remember(*out, cost);
  VLOG(VL_VERBOSE, "graph exec %p: %lld tokens per launch", (void *)*out, cost);
}

VGPU_EXPORT CUresult cuGraphInstantiateWithFlags(CUgraphExec *out, CUgraph g, unsigned long long flags) {
  vgpu_boot();
  if (unlikely(!R.cuGraphInstantiateWithFlags)) return CUDA_ERROR_NOT_FOUND;
  CUresult r = R.cuGraphInstantiateWithFlags(out, g, flags);
  graph_instantiated(out, g, r);
  return r;
}
VGPU_EXPORT CUresult cuGraphInstantiateWithParams(CUgraphExec *out, CUgraph g, void *params) {
  vgpu_boot();
  if (unlikely(!R.cuGraphInstantiateWithParams)) return CUDA_ERROR_NOT_FOUND;
  CUresult r = R.cuGraphInstantiateWithParams(out, g, params);
  graph_instantiated(out, g, r);
  return r;
}
VGPU_EXPORT CUresult cuGraphInstantiateWithParams_ptsz(CUgraphExec *out, CUgraph g, void *params) {
  vgpu_boot();
  if (unlikely(!R.cuGraphInstantiateWithParams_ptsz)) return CUDA_ERROR_NOT_FOUND;
  CUresult r = R.cuGraphInstantiateWithParams_ptsz(out, g, params);
  graph_instantiated(out, g, r);
  return r;
}
VGPU_EXPORT CUresult cuGraphExecDestroy(CUgraphExec exec) {
  vgpu_boot();
  if (vgpu_graph_limit_enabled()) graph_lookup(exec, NULL, 1);
  return R.cuGraphExecDestroy ? R.cuGraphExecDestroy(exec) : CUDA_ERROR_NOT_FOUND;
}

static inline unsigned graph_tokens(CUgraphExec exec) {
  long long cost = 0;
  if (!vgpu_graph_limit_enabled() || !graph_lookup(exec, &cost, 0) || cost <= 0) return 0;
  return cost > 0x7fffffffll ? 0x7fffffffu : (unsigned)cost;
}
VGPU_EXPORT CUresult cuGraphLaunch(CUgraphExec exec, CUstream s) {
  if (unlikely(!G_cfg)) vgpu_boot();
  if (unlikely(!R.cuGraphLaunch)) return CUDA_ERROR_NOT_FOUND;
  unsigned tokens = graph_tokens(exec);
  if (!tokens) return R.cuGraphLaunch(exec, s);
  LIMITED_LAUNCH(tokens, 1, 1, s, 0, R.cuGraphLaunch(exec, s));
}
VGPU_EXPORT CUresult cuGraphLaunch_ptsz(CUgraphExec exec, CUstream s) {
  if (unlikely(!G_cfg)) vgpu_boot();
  if (unlikely(!R.cuGraphLaunch_ptsz)) return CUDA_ERROR_NOT_FOUND;
  unsigned tokens = graph_tokens(exec);
  if (!tokens) return R.cuGraphLaunch_ptsz(exec, s);
  LIMITED_LAUNCH(tokens, 1, 1, s, 1, R.cuGraphLaunch_ptsz(exec, s));
}
