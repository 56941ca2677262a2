/*
 * memgate.c - memory-cap / oversubscription hooks.
 *
 * What stays on the host: calling NVML for the per-process lists (the driver is the only
 * source), the container-membership file reads, the cross-process locks and the calls into
 * the real allocator.  What moved to the device: the fold of those lists into the container's
 * `used`, the ledger sum, the quota check, the GPU/UVA/OOM decision and the reported
 * total/used/free numbers (vgpu_quota_kernel), plus the record table of UVA allocations
 * (vgpu_slab_*_kernel).  If the device runtime cannot be brought up in a process that HAS a CUDA
 * context the hooks fail loudly (CUDA_ERROR_NOT_SUPPORTED): allocation decisions have no CPU
 * path.  The one host evaluation is the NVML memory report for processes without any CUDA
 * context (nvidia-smi and friends), which the reference also serves (nvml_hook.c:47-103).
 *
 * Behavioural contract (reference library/src/cuda_hook.c):
 *   prepare_memory_allocation :93-116      load_limited_memory_view :118-136
 *   get_used_gpu_memory       :807-920     allocation hooks        :1316-1697, :2004-2050
 *   cuDeviceTotalMem/MemGetInfo :1699-1808 free hooks              :2052-2109
 * and library/src/loader.c:1824-1922 (UVA ledger), library/src/nvml_hook.c:47-126.
 */
#include "vgpu_internal.h"

#include <time.h>

/* ------------------------------------------------------------------ phase profile (VGPU_B200_PROFILE=1)
 * Where the time of an intercepted allocation / free goes, summed per phase and printed at exit:
 * development aid for the allocator-storm comparison against the reference. */
enum { PH_LOCK, PH_ARM, PH_COMPUTE_LIST, PH_STAGE, PH_GRAPHICS_LIST, PH_COLLECT, PH_DRIVER_ALLOC, PH_UNLOCK, PH_FREE_PRE, PH_DRIVER_FREE,
       PH_FREE_POST, PH_COUNT };
static const char *g_ph_names[PH_COUNT] = {"lock", "arm", "nvml_compute_list", "stage", "nvml_graphics_list", "collect", "driver_alloc",
                                           "unlock", "free_pre", "driver_free", "free_post"};
static uint64_t g_ph_ns[PH_COUNT], g_ph_n[PH_COUNT];
static int g_prof = -1;
static inline uint64_t prof_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
static void prof_dump(void) {
  for (int i = 0; i < PH_COUNT; i++)
    if (g_ph_n[i]) fprintf(stderr, "[vGPU PROFILE] %-20s n %8llu  mean %8.0f ns\n", g_ph_names[i], (unsigned long long)g_ph_n[i], (double)g_ph_ns[i] / (double)g_ph_n[i]);
}
static inline int prof_on(void) {
  if (unlikely(g_prof < 0)) {
    const char *e = getenv("VGPU_B200_PROFILE");
    g_prof = e && *e == '1';
    if (g_prof) atexit(prof_dump);
  }
  return g_prof;
}
#define PROF_T0() uint64_t pt_ = prof_on() ? prof_now() : 0
#define PROF(ph)                                 \
  do {                                           \
    if (g_prof > 0) {                            \
      uint64_t n_ = prof_now();                  \
      g_ph_ns[ph] += n_ - pt_;                   \
      g_ph_n[ph]++;                              \
      pt_ = n_;                                  \
    }                                            \
  } while (0)

/* ------------------------------------------------------------------ limited memory view */
typedef struct {
  int host_index;
  int lock_fd;
  int limited;      /* the device has a memory cap and the view was computed */
  int failed;       /* the device runtime is unavailable                     */
  vgpu_dev_rt *rt;
  vgpu_quota_res_t res;
} memview_t;

static uint32_t fetch_list(nvmlDevice_t nv, int graphics, vgpu_proc_t *out, nvmlReturn_t *rc) {
  unsigned int n = VGPU_MAX_PIDS;
  nvmlReturn_t r;
  nvmlReturn_t (*v1)(nvmlDevice_t, unsigned int *, vgpu_proc_t *) =
      graphics ? R.nvmlDeviceGetGraphicsRunningProcesses : R.nvmlDeviceGetComputeRunningProcesses;
  nvmlReturn_t (*v3)(nvmlDevice_t, unsigned int *, vgpu_proc_v2_t *) =
      graphics ? R.nvmlDeviceGetGraphicsRunningProcesses_v3 : R.nvmlDeviceGetComputeRunningProcesses_v3;
  if (v1) {
    r = v1(nv, &n, out); /* unversioned symbol == 16-byte v1 records (nvml-subset.h:81-88) */
  } else if (v3) {
    /* the reference would hand its 16-byte array to the 24-byte ABI here; convert instead */
    static __thread vgpu_proc_v2_t wide[VGPU_MAX_PIDS];
    r = v3(nv, &n, wide);
    if (r == NVML_SUCCESS)
      for (unsigned int i = 0; i < n; i++) {
        out[i].pid = wide[i].pid;
        out[i]._pad = 0;
        out[i].used_bytes = wide[i].used_bytes;
      }
  } else {
    r = NVML_ERROR_FUNCTION_NOT_FOUND;
  }
  *rc = r;
  return r == NVML_SUCCESS ? n : 0;
}

/* Fill the request block with everything the kernel needs to restate get_used_gpu_memory +
 * get_used_gpu_virt_memory - except the graphics list (stage_graphics).  Caller holds the
 * per-GPU file lock (and rt->q_mu when `q` is a runtime's block).  Returns 0 when the compute
 * list is unavailable: used = 0 and the graphics list is not consulted (:825-830). */
static int stage_request(vgpu_quota_req_t *q, int host_index, nvmlDevice_t nv) {
  const vgpu_cfg_dev_t *c = &G_cfg->devices[host_index];
  q->mode = (uint32_t)G_cfg->compatibility_mode;
  q->memory_oversold = (uint32_t)c->memory_oversold;
  q->total_memory = c->total_memory;
  q->real_memory = c->real_memory;
  q->self_pid = (uint32_t)getpid();
  q->n_compute = q->n_vmem = 0;
  int have_compute = 0;
  if (nv) {
    nvmlReturn_t rc;
    uint32_t nc = fetch_list(nv, 0, q->compute, &rc);
    if (rc != NVML_SUCCESS) {
      VLOG(VL_ERROR, "nvmlDeviceGetComputeRunningProcesses call failed, return: %d, str: %s", rc,
           vgpu_nv_err(rc));
    } else {
      have_compute = 1;
      q->n_compute = nc;
      if (G_cfg->compatibility_mode != VGPU_MODE_HOST) {
        static __thread uint32_t pids[VGPU_MAX_PIDS];
        for (uint32_t i = 0; i < nc; i++) pids[i] = q->compute[i].pid;
        vgpu_pid_flags(pids, nc, q->cflags);
      }
    }
  }
  if (G_cfg->vmem_node && G_vmem) {
    int fd = vgpu_vmem_lock(host_index, 0);
    if (fd >= 0) {
      const vgpu_vmem_dev_t *d = &G_vmem->devices[host_index];
      uint32_t n = d->processes_size > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : d->processes_size;
      memcpy(q->vmem, d->processes, (size_t)n * sizeof(vgpu_vmem_rec_t));
      q->n_vmem = n;
      vgpu_vmem_unlock(fd, host_index);
    }
  }
  return have_compute;
}

/* The graphics list, fetched into a scratch copy first: the allocation path evaluates the
 * request speculatively against the previous call's graphics list (still in the block) while
 * this ioctl is in flight, and only re-evaluates when the list really changed. */
typedef struct {
  uint32_t n;
  vgpu_proc_t recs[VGPU_MAX_PIDS];
  uint8_t flags[VGPU_MAX_PIDS];
} gfx_scratch_t;

static void fetch_graphics(gfx_scratch_t *g, nvmlDevice_t nv) {
  nvmlReturn_t rc;
  g->n = fetch_list(nv, 1, g->recs, &rc);
  if (rc != NVML_SUCCESS) {
    VLOG(VL_ERROR, "nvmlDeviceGetGraphicsRunningProcesses call failed, return: %d, str: %s", rc,
         vgpu_nv_err(rc));
    g->n = 0;
  }
  if (g->n && G_cfg->compatibility_mode != VGPU_MODE_HOST) {
    static __thread uint32_t pids[VGPU_MAX_PIDS];
    for (uint32_t i = 0; i < g->n; i++) pids[i] = g->recs[i].pid;
    vgpu_pid_flags(pids, g->n, g->flags);
  } else {
    memset(g->flags, 0, g->n);
  }
}

static int graphics_same(const vgpu_quota_req_t *q, const gfx_scratch_t *g) {
  return q->n_graphics == g->n && (g->n == 0 || (memcmp(q->graphics, g->recs, (size_t)g->n * sizeof(vgpu_proc_t)) == 0 &&
                                                 memcmp(q->gflags, g->flags, g->n) == 0));
}

static void graphics_install(vgpu_quota_req_t *q, const gfx_scratch_t *g) {
  memcpy(q->graphics, g->recs, (size_t)g->n * sizeof(vgpu_proc_t));
  memcpy(q->gflags, g->flags, g->n);
  q->n_graphics = g->n;
}

extern int vgpu_rt_quota(vgpu_dev_rt *rt, vgpu_quota_res_t *out);
extern int vgpu_rt_slab_insert(vgpu_dev_rt *rt, CUdeviceptr dptr, uint64_t bytes);
extern int vgpu_rt_slab_remove(vgpu_dev_rt *rt, CUdeviceptr dptr, uint64_t *bytes);

/* ------------------------------------------------------------------ context-less evaluation
 * An NVML-only client (nvidia-smi, pynvml, a metrics exporter, a framework probing memory
 * before it forks its CUDA workers) has no CUDA context for vgpu_quota_kernel to run in.  The
 * reference answers those on the CPU (nvml_hook.c:47-103), so does this: the same fold as the
 * kernel (membership latch cuda_hook.c:735-805, graphics/compute dedup :868-887, ledger sum
 * loader.c:1909-1922, clamp nvml_hook.c:58-63), written for one thread.  Only the NVML report
 * uses it; allocation decisions always run on the device. */
static uint64_t host_fold(uint32_t mode, const vgpu_proc_t *p, const uint8_t *flags, const uint8_t *dead, uint32_t n) {
  int open_mode = (mode & VGPU_MODE_OPEN_KERNEL) == VGPU_MODE_OPEN_KERNEL;
  int ladder = (mode & VGPU_MODE_CLIENT) == VGPU_MODE_CLIENT || (mode & VGPU_MODE_CGROUPV2) == VGPU_MODE_CGROUPV2 ||
               (mode & VGPU_MODE_CGROUPV1) == VGPU_MODE_CGROUPV1;
  uint64_t sum = 0;
  int regime = 0; /* 0 undecided, 1 primary, 2 open-kernel */
  for (uint32_t i = 0; i < n; i++) {
    if (dead && dead[i]) continue;
    int prim = (flags[i] & VGPU_FLAG_PRIMARY) != 0, loc = open_mode && (flags[i] & VGPU_FLAG_LOCAL);
    if (ladder) {
      if (regime != 2 && prim) { regime = 1; sum += p[i].used_bytes; }
      else if (regime != 1 && loc) { regime = 2; sum += p[i].used_bytes; }
    } else if (open_mode) {
      if (flags[i] & VGPU_FLAG_LOCAL) sum += p[i].used_bytes;
    } else if (mode == VGPU_MODE_HOST) {
      sum += p[i].used_bytes;
    }
  }
  return sum;
}

static void host_nvml_view(const vgpu_quota_req_t *q, vgpu_quota_res_t *r) {
  static __thread uint8_t dead[VGPU_MAX_PIDS];
  uint32_t nc = q->n_compute > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : q->n_compute;
  uint32_t ng = q->n_graphics > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : q->n_graphics;
  uint32_t nv = q->n_vmem > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : q->n_vmem;
  uint64_t used = host_fold(q->mode, q->compute, q->cflags, NULL, nc);
  for (uint32_t i = 0; i < ng; i++) { /* graphics pids already in the compute list are dropped */
    dead[i] = 0;
    for (uint32_t j = 0; j < nc && !dead[i]; j++) dead[i] = q->compute[j].pid == q->graphics[i].pid;
  }
  used += host_fold(q->mode, q->graphics, q->gflags, dead, ng);
  /* footprint of the container's CUDA-holding siblings that run this library (none of ours:
   * this process has no context) */
  used = used >= q->self_bytes ? used - q->self_bytes : 0;
  uint64_t vmem = 0;
  for (uint32_t i = 0; i < nv; i++) vmem += q->vmem[i].used;
  uint64_t tu = used + vmem;
  memset(r, 0, sizeof *r);
  r->used = used;
  r->vmem = vmem;
  r->total = q->total_memory;
  r->out_used = tu >= q->total_memory ? q->total_memory : tu;
  r->out_free = r->total - r->out_used;
}

/* Evaluate (kind, request) for one GPU.  On return the per-GPU lock is still held
 * (mv->lock_fd) exactly like load_limited_memory_view leaves it. */
static void memview(memview_t *mv, CUdevice dev, int host_index, nvmlDevice_t nv, uint32_t kind,
                    uint64_t request, int allow_uva, int real_ok, uint64_t real_total) {
  memset(mv, 0, sizeof *mv);
  mv->lock_fd = -1;
  mv->host_index = host_index;
  if (host_index < 0 || !G_cfg->devices[host_index].memory_limit) return;
  vgpu_dev_rt *rt = vgpu_rt_get(host_index, dev);
  if (!rt) {
    mv->failed = 1;
    return;
  }
  mv->rt = rt;
  PROF_T0();
  mv->lock_fd = vgpu_lock_gpu(host_index);
  PROF(PH_LOCK);
  if (!nv) nv = vgpu_nvml_handle_of_host(host_index);
  if (!nv) VLOG(VL_ERROR, "cuda device %d cannot find the corresponding nvml devices", dev);
  pthread_mutex_lock(&rt->q_mu);
  vgpu_quota_req_t *q = rt->q_req;
  /* launch first: the kernel waits on the device for the publication while the host talks to NVML */
  uint32_t armed = (rt->quota_armed && nv) ? vgpu_rt_quota_arm(rt) : 0;
  PROF(PH_ARM);
  int have_compute = stage_request(q, host_index, nv);
  PROF(PH_COMPUTE_LIST);
  /* footprint of every live library instance of this container on this GPU (lock is held) */
  q->self_bytes = mv->lock_fd >= 0 ? vgpu_self_registry(host_index, rt->self_bytes, 0) : rt->self_bytes;
  q->kind = kind;
  q->request = request;
  q->allow_uva = (uint32_t)allow_uva;
  q->real_ok = (uint32_t)real_ok;
  q->real_total = real_total;
  int done = 0;
  if (!have_compute) {
    q->n_graphics = 0;
    rt->gfx_valid = 0;
    if (armed) vgpu_rt_quota_publish(rt, armed);
  } else {
    static __thread gfx_scratch_t scratch;
    if (armed && rt->gfx_valid) {
      /* speculate: graphics list unchanged since the previous evaluation (it is empty on a
       * compute-only part); the kernel decides while the second ioctl is in flight */
      vgpu_rt_quota_publish(rt, armed);
      PROF(PH_STAGE);
      fetch_graphics(&scratch, nv);
      PROF(PH_GRAPHICS_LIST);
      if (graphics_same(q, &scratch)) {
        done = vgpu_rt_quota_collect(rt, armed, &mv->res) == 0;
        PROF(PH_COLLECT);
      } else {
        vgpu_quota_res_t drop;
        vgpu_rt_quota_collect(rt, armed, &drop); /* let the speculative evaluation finish before restaging */
        graphics_install(q, &scratch);
      }
      armed = 0;
    } else {
      fetch_graphics(&scratch, nv);
      graphics_install(q, &scratch);
      rt->gfx_valid = 1;
      if (armed) vgpu_rt_quota_publish(rt, armed);
    }
  }
  if (!done && armed) done = vgpu_rt_quota_collect(rt, armed, &mv->res) == 0;
  if (!done) {
    rt->q_req_self_set = 1;
    if (vgpu_rt_quota(rt, &mv->res)) mv->failed = 1;
    else done = 1;
  }
  if (done) mv->limited = 1;
  pthread_mutex_unlock(&rt->q_mu);
}

/* ------------------------------------------------------------------ UVA ledger (host file + device slab) */
static void ledger_add(vgpu_dev_rt *rt, CUdeviceptr dptr, uint64_t bytes, int host_index) {
  if (rt && vgpu_rt_slab_insert(rt, dptr, bytes))
    VLOG(VL_ERROR, "failed to record virt memory node (slab full)");
  if (host_index < 0 || host_index >= VGPU_MAX_DEVICES || !G_vmem) return;
  int fd = vgpu_vmem_lock(host_index, 1);
  if (fd < 0) return;
  vgpu_vmem_dev_t *d = &G_vmem->devices[host_index];
  int me = getpid();
  uint32_t n = d->processes_size, i;
  for (i = 0; i < n; i++)
    if (d->processes[i].pid == me) {
      d->processes[i].used += bytes;
      break;
    }
  if (i == n) {
    if (n >= VGPU_MAX_PIDS) {
      VLOG(VL_ERROR, "host device %d virtual memory process list is full", host_index);
    } else {
      d->processes[n].pid = me;
      d->processes[n].used = bytes;
      d->processes_size = n + 1;
    }
  }
  vgpu_vmem_unlock(fd, host_index);
}

static void ledger_sub(CUdevice dev, CUdeviceptr dptr) {
  int host_index = vgpu_host_index_of_cuda(dev);
  int slot = host_index >= 0 ? host_index : dev;
  uint64_t bytes = 0;
  int found = 0;
  vgpu_dev_rt *rt = vgpu_rt_peek(slot);
  if (rt && vgpu_rt_slab_remove(rt, dptr, &bytes) == 0) found = 1;
  /* The reference keeps ONE list of UVA allocations per process and subtracts a freed node from the ledger record of
   * whatever device is current at the free (loader.c:1869-1907) - also when the allocation was recorded under another
   * device.  The records live in per-device tables here, so a miss on the current device's table is followed by the
   * other runtimes' (skipped without a launch while they hold no record). */
  for (int h = 0; !found && h < VGPU_MAX_DEVICES; h++) {
    vgpu_dev_rt *other = vgpu_rt_peek(h);
    if (!other || other == rt) continue;
    if (vgpu_rt_slab_remove(other, dptr, &bytes) == 0) found = 1;
  }
  /* ... and by the records whose context is gone (the reference's list survives a device reset) */
  if (!found && vgpu_stale_uva_remove(dptr, &bytes) == 0) found = 1;
  if (!found) return;
  if (host_index < 0 || host_index >= VGPU_MAX_DEVICES || !G_vmem) return;
  int fd = vgpu_vmem_lock(host_index, 1);
  if (fd < 0) return;
  vgpu_vmem_dev_t *d = &G_vmem->devices[host_index];
  int me = getpid();
  for (uint32_t i = 0; i < d->processes_size; i++)
    if (d->processes[i].pid == me) {
      d->processes[i].used = d->processes[i].used >= bytes ? d->processes[i].used - bytes : 0;
      break;
    }
  vgpu_vmem_unlock(fd, host_index);
}

/* ------------------------------------------------------------------ allocation front half */
typedef struct {
  CUresult early;   /* != SUCCESS: return this without touching the driver */
  int path;         /* VGPU_PATH_* */
  memview_t mv;
  CUdevice dev;
} gate_t;

static void gate_open(gate_t *g, uint64_t request, int allow_uva, const vcu_mem_alloc_prop_t *prop) {
  memset(g, 0, sizeof *g);
  g->mv.lock_fd = -1;
  g->mv.host_index = -1;
  g->path = VGPU_PATH_GPU;
  if (unlikely(!G_cfg)) vgpu_boot();
  CUresult r = R.cuCtxGetDevice ? R.cuCtxGetDevice(&g->dev) : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) {
    if (prop && prop->location.type == 1) g->dev = prop->location.id; /* cuMemCreate w/o ctx (:1678) */
    else { g->early = r; return; }
  }
  int host_index = vgpu_host_index_of_cuda(g->dev);
  memview(&g->mv, g->dev, host_index, NULL, VGPU_Q_ALLOC, request, allow_uva, 0, 0);
  if (g->mv.failed) {
    g->early = CUDA_ERROR_NOT_SUPPORTED;
    return;
  }
  if (g->mv.limited) g->path = (int)g->mv.res.path;
  if (g->path == VGPU_PATH_OOM) {
    vgpu_metric_add(host_index, VM_OOM_LIMIT, 1);
    g->early = CUDA_ERROR_OUT_OF_MEMORY;
  }
}

static inline void gate_close(gate_t *g) { vgpu_unlock_gpu(g->mv.lock_fd); }

static int oversold(const gate_t *g) {
  return g->mv.host_index >= 0 && G_cfg->devices[g->mv.host_index].memory_oversold;
}

/* GPU path failed with driver OOM on an oversold device => retry through UVA (:1372-1386) */
static CUresult to_uva(gate_t *g, CUdeviceptr *dptr, size_t bytes) {
  CUresult r = R.cuMemAllocManaged ? R.cuMemAllocManaged(dptr, bytes, VCU_MEM_ATTACH_GLOBAL)
                                   : CUDA_ERROR_NOT_FOUND;
  VLOG(VL_VERBOSE, "cuMemAllocManaged to allocate unified memory (oversold), size: %zu, ret: %d", bytes, r);
  if (r == CUDA_SUCCESS) {
    vgpu_dev_rt *rt = g->mv.rt ? g->mv.rt : vgpu_rt_get(g->mv.host_index, g->dev);
    ledger_add(rt, *dptr, bytes, g->mv.host_index);
  }
  return r;
}

#define DRIVER_OOM_RETRY(g, r)                                               \
  ((r) == CUDA_ERROR_OUT_OF_MEMORY && oversold(g) &&                         \
   (vgpu_metric_add((g)->mv.host_index, VM_OOM_DRIVER, 1),                   \
    vgpu_metric_add((g)->mv.host_index, VM_UVA_FALLBACK, 1), 1))

VGPU_EXPORT CUresult cuMemAllocManaged(CUdeviceptr *dptr, size_t bytes, unsigned int flags) {
  gate_t g;
  gate_open(&g, bytes, 1, NULL);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS) {
    if (g.path == VGPU_PATH_UVA) flags = VCU_MEM_ATTACH_GLOBAL;
    r = R.cuMemAllocManaged ? R.cuMemAllocManaged(dptr, bytes, flags) : CUDA_ERROR_NOT_FOUND;
    /* every GLOBAL-attached managed allocation is ledgered, requested or rerouted (:1336) */
    if (r == CUDA_SUCCESS && flags == VCU_MEM_ATTACH_GLOBAL) {
      vgpu_dev_rt *rt = g.mv.rt ? g.mv.rt : vgpu_rt_get(g.mv.host_index, g.dev);
      ledger_add(rt, *dptr, bytes, g.mv.host_index);
    }
  }
  gate_close(&g);
  return r;
}

static CUresult alloc_linear(CUdeviceptr *dptr, size_t bytes) {
  gate_t g;
  gate_open(&g, bytes, 1, NULL);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS && g.mv.limited && oversold(&g) && vgpu_slab_mode()) {
    /* VGPU_B200_SLAB=1: same decision, same accounting, but the library owns the placement -
     * a "UVA" decision spills the coldest slab instead of leaving it to UVM (slabmode.c) */
    int recorded = 0;
    CUresult sr = vgpu_slab_alloc(g.mv.rt, g.dev, g.path, dptr, bytes, &recorded);
    if (sr != CUDA_ERROR_NOT_SUPPORTED) {
      if (sr == CUDA_SUCCESS && recorded) {
        if (g.path != VGPU_PATH_UVA) { /* the driver ran out of HBM on the GPU path (:1372-1386) */
          vgpu_metric_add(g.mv.host_index, VM_OOM_DRIVER, 1);
          vgpu_metric_add(g.mv.host_index, VM_UVA_FALLBACK, 1);
        }
        ledger_add(g.mv.rt, *dptr, bytes, g.mv.host_index);
      }
      gate_close(&g);
      return sr;
    }
  }
  if (r == CUDA_SUCCESS) {
    if (g.path == VGPU_PATH_UVA) {
      r = to_uva(&g, dptr, bytes);
    } else {
      PROF_T0();
      r = R.cuMemAlloc_v2 ? R.cuMemAlloc_v2(dptr, bytes)
          : R.cuMemAlloc  ? R.cuMemAlloc(dptr, bytes)
                          : CUDA_ERROR_NOT_FOUND;
      PROF(PH_DRIVER_ALLOC);
      if (DRIVER_OOM_RETRY(&g, r)) r = to_uva(&g, dptr, bytes);
    }
  }
  {
    PROF_T0();
    gate_close(&g);
    PROF(PH_UNLOCK);
  }
  return r;
}
VGPU_EXPORT CUresult cuMemAlloc_v2(CUdeviceptr *dptr, size_t bytes) { return alloc_linear(dptr, bytes); }
VGPU_EXPORT CUresult cuMemAlloc(CUdeviceptr *dptr, size_t bytes) { return alloc_linear(dptr, bytes); }

static CUresult alloc_pitch(CUdeviceptr *dptr, size_t *pitch, size_t width, size_t height, unsigned elem) {
  /* request = guessed pitch x height (:1411-1412) */
  size_t guess = (((width - 1) / elem) + 1) * elem;
  size_t request = guess * height;
  gate_t g;
  gate_open(&g, request, 1, NULL);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS) {
    int uva = g.path == VGPU_PATH_UVA;
    if (!uva) {
      r = R.cuMemAllocPitch_v2 ? R.cuMemAllocPitch_v2(dptr, pitch, width, height, elem)
          : R.cuMemAllocPitch  ? R.cuMemAllocPitch(dptr, pitch, width, height, elem)
                               : CUDA_ERROR_NOT_FOUND;
      uva = DRIVER_OOM_RETRY(&g, r);
    }
    if (uva) {
      r = to_uva(&g, dptr, request);
      if (r == CUDA_SUCCESS) *pitch = guess;
    }
  }
  gate_close(&g);
  return r;
}
VGPU_EXPORT CUresult cuMemAllocPitch_v2(CUdeviceptr *d, size_t *p, size_t w, size_t h, unsigned e) {
  return alloc_pitch(d, p, w, h, e);
}
VGPU_EXPORT CUresult cuMemAllocPitch(CUdeviceptr *d, size_t *p, size_t w, size_t h, unsigned e) {
  return alloc_pitch(d, p, w, h, e);
}

static CUresult alloc_async(CUdeviceptr *dptr, size_t bytes, CUstream s, int ptsz) {
  gate_t g;
  gate_open(&g, bytes, 1, NULL);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS) {
    if (g.path == VGPU_PATH_UVA) {
      r = to_uva(&g, dptr, bytes);
    } else {
      CUresult (*fn)(CUdeviceptr *, size_t, CUstream) = ptsz ? R.cuMemAllocAsync_ptsz : R.cuMemAllocAsync;
      r = fn ? fn(dptr, bytes, s) : CUDA_ERROR_NOT_FOUND;
      if (DRIVER_OOM_RETRY(&g, r)) r = to_uva(&g, dptr, bytes);
    }
  }
  gate_close(&g);
  return r;
}
VGPU_EXPORT CUresult cuMemAllocAsync(CUdeviceptr *d, size_t n, CUstream s) { return alloc_async(d, n, s, 0); }
VGPU_EXPORT CUresult cuMemAllocAsync_ptsz(CUdeviceptr *d, size_t n, CUstream s) { return alloc_async(d, n, s, 1); }

/* the cap applies but these can never spill (allow_uva = 0, Appendix B.15) */
static CUresult alloc_pool(CUdeviceptr *dptr, size_t bytes, CUmemoryPool pool, CUstream s, int ptsz) {
  gate_t g;
  gate_open(&g, bytes, 0, NULL);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS) {
    CUresult (*fn)(CUdeviceptr *, size_t, CUmemoryPool, CUstream) =
        ptsz ? R.cuMemAllocFromPoolAsync_ptsz : R.cuMemAllocFromPoolAsync;
    r = fn ? fn(dptr, bytes, pool, s) : CUDA_ERROR_NOT_FOUND;
  }
  gate_close(&g);
  return r;
}
VGPU_EXPORT CUresult cuMemAllocFromPoolAsync(CUdeviceptr *d, size_t n, CUmemoryPool p, CUstream s) {
  return alloc_pool(d, n, p, s, 0);
}
VGPU_EXPORT CUresult cuMemAllocFromPoolAsync_ptsz(CUdeviceptr *d, size_t n, CUmemoryPool p, CUstream s) {
  return alloc_pool(d, n, p, s, 1);
}

VGPU_EXPORT CUresult cuMemCreate(CUmemGenericAllocationHandle *h, size_t size,
                                 const vcu_mem_alloc_prop_t *prop, unsigned long long flags) {
  gate_t g;
  gate_open(&g, size, 0, prop);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS) r = R.cuMemCreate ? R.cuMemCreate(h, size, prop, flags) : CUDA_ERROR_NOT_FOUND;
  gate_close(&g);
  return r;
}

/* array requests are sized in *bits* per channel - a reference quirk we keep (:1546-1569) */
static size_t array_unit(int format) {
  switch (format) {
  case 0x01: case 0x08: return 8;
  case 0x02: case 0x09: case 0x10: return 16;
  default: return 32;
  }
}

static CUresult array2d(CUarray *h, const vcu_array_desc_t *d) {
  gate_t g;
  gate_open(&g, array_unit(d->Format) * d->NumChannels * d->Height * d->Width, 0, NULL);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS)
    r = R.cuArrayCreate_v2 ? R.cuArrayCreate_v2(h, d) : R.cuArrayCreate ? R.cuArrayCreate(h, d) : CUDA_ERROR_NOT_FOUND;
  gate_close(&g);
  return r;
}
VGPU_EXPORT CUresult cuArrayCreate_v2(CUarray *h, const vcu_array_desc_t *d) { return array2d(h, d); }
VGPU_EXPORT CUresult cuArrayCreate(CUarray *h, const vcu_array_desc_t *d) { return array2d(h, d); }

static size_t array3d_request(const vcu_array3d_desc_t *d) {
  return array_unit(d->Format) * d->NumChannels * d->Height * d->Width * d->Depth;
}
static CUresult array3d(CUarray *h, const vcu_array3d_desc_t *d) {
  gate_t g;
  gate_open(&g, array3d_request(d), 0, NULL);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS)
    r = R.cuArray3DCreate_v2 ? R.cuArray3DCreate_v2(h, d)
        : R.cuArray3DCreate  ? R.cuArray3DCreate(h, d)
                             : CUDA_ERROR_NOT_FOUND;
  gate_close(&g);
  return r;
}
VGPU_EXPORT CUresult cuArray3DCreate_v2(CUarray *h, const vcu_array3d_desc_t *d) { return array3d(h, d); }
VGPU_EXPORT CUresult cuArray3DCreate(CUarray *h, const vcu_array3d_desc_t *d) { return array3d(h, d); }

VGPU_EXPORT CUresult cuMipmappedArrayCreate(CUmipmappedArray *h, const vcu_array3d_desc_t *d,
                                            unsigned int levels) {
  gate_t g;
  gate_open(&g, array3d_request(d), 0, NULL);
  CUresult r = g.early;
  if (r == CUDA_SUCCESS)
    r = R.cuMipmappedArrayCreate ? R.cuMipmappedArrayCreate(h, d, levels) : CUDA_ERROR_NOT_FOUND;
  gate_close(&g);
  return r;
}

/* ------------------------------------------------------------------ free */
/* Tenant isolation option (VGPU_B200_SCRUB_ON_FREE=1): zero an allocation with the 128-bit clear
 * kernel before it goes back to the driver, so the next tenant of the shared GPU cannot read it.
 * The reference has no counterpart (it never touches tenant memory); accounting is unaffected. */
static int scrub_on_free(void) {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("VGPU_B200_SCRUB_ON_FREE");
    on = e && (!strcmp(e, "1") || !strcmp(e, "true"));
  }
  return on;
}

static void scrub(vgpu_dev_rt *rt, CUdeviceptr dptr) {
  CUdeviceptr base = 0;
  size_t size = 0;
  if (!R.cuMemGetAddressRange_v2 || R.cuMemGetAddressRange_v2(&base, &size, dptr) != CUDA_SUCCESS || !size) return;
  if (R.cuCtxSynchronize) R.cuCtxSynchronize(); /* cuMemFree would wait for pending work anyway */
  if (vgpu_rt_clear(rt, base, size, rt->q_stream) == CUDA_SUCCESS && R.cuStreamSynchronize(rt->q_stream) == CUDA_SUCCESS)
    vgpu_metric_add(rt->host_index, VM_SCRUBBED_BYTES, size);
}

static CUresult free_sync(CUdeviceptr dptr) {
  CUdevice dev;
  if (unlikely(!G_cfg)) vgpu_boot();
  CUresult r = R.cuCtxGetDevice ? R.cuCtxGetDevice(&dev) : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) return r;
  /* cuMemFree synchronises the device: it must not wait for the resident governor */
  vgpu_dev_rt *rt = vgpu_rt_peek(vgpu_host_index_of_cuda(dev));
  if (vgpu_slab_mode()) { /* a slab of the VGPU_B200_SLAB mode is not the driver's to free */
    int was_uva = 0;
    uint64_t sbytes = 0;
    /* the current device's table first; a pointer may also be freed while another device is current (it is
     * a unified address, the driver does not care) - then the slab belongs to one of the other runtimes.  The
     * ledger is only touched for the current device, like the reference (loader.c:1875-1907 searches the
     * list of the host index it is given) */
    if (rt && rt->vs_host && vgpu_slab_free(rt, dptr, &r, &was_uva, &sbytes)) {
      if (r == CUDA_SUCCESS) ledger_sub(dev, dptr);
      return r;
    }
    for (int h = 0; h < VGPU_MAX_DEVICES; h++) {
      vgpu_dev_rt *other = vgpu_rt_peek(h);
      if (!other || other == rt || !other->vs_host) continue;
      if (vgpu_slab_free(other, dptr, &r, &was_uva, &sbytes)) {
        if (r == CUDA_SUCCESS) ledger_sub(dev, dptr);
        return r;
      }
    }
  }
  PROF_T0();
  if (rt) vgpu_limiter_before_blocking_call(rt);
  if (rt) vgpu_limiter_quiesce(rt);
  if (rt && scrub_on_free()) scrub(rt, dptr);
  PROF(PH_FREE_PRE);
  r = R.cuMemFree_v2 ? R.cuMemFree_v2(dptr) : R.cuMemFree ? R.cuMemFree(dptr) : CUDA_ERROR_NOT_FOUND;
  PROF(PH_DRIVER_FREE);
  if (rt) vgpu_limiter_resume(rt, 0);
  if (r == CUDA_SUCCESS) ledger_sub(dev, dptr);
  PROF(PH_FREE_POST);
  return r;
}
VGPU_EXPORT CUresult cuMemFree_v2(CUdeviceptr p) { return free_sync(p); }
VGPU_EXPORT CUresult cuMemFree(CUdeviceptr p) { return free_sync(p); }

static CUresult free_async(CUdeviceptr dptr, CUstream s, int ptsz) {
  CUdevice dev;
  if (unlikely(!G_cfg)) vgpu_boot();
  CUresult r = R.cuCtxGetDevice ? R.cuCtxGetDevice(&dev) : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) return r;
  if (vgpu_slab_mode()) /* slabs are remapped, not pooled: freed synchronously (whichever device owns them) */
    for (int h = 0; h < VGPU_MAX_DEVICES; h++) {
      vgpu_dev_rt *rt = vgpu_rt_peek(h);
      if (rt && rt->vs_host) return free_sync(dptr);
    }
  CUresult (*fn)(CUdeviceptr, CUstream) = ptsz ? R.cuMemFreeAsync_ptsz : R.cuMemFreeAsync;
  r = fn ? fn(dptr, s) : CUDA_ERROR_NOT_FOUND;
  if (r == CUDA_SUCCESS) ledger_sub(dev, dptr);
  return r;
}
VGPU_EXPORT CUresult cuMemFreeAsync(CUdeviceptr p, CUstream s) { return free_async(p, s, 0); }
VGPU_EXPORT CUresult cuMemFreeAsync_ptsz(CUdeviceptr p, CUstream s) { return free_async(p, s, 1); }

/* ------------------------------------------------------------------ reported sizes */
static CUresult total_mem(size_t *bytes, CUdevice dev) {
  if (unlikely(!G_cfg)) vgpu_boot();
  int h = vgpu_host_index_of_cuda(dev);
  if (h >= 0 && G_cfg->devices[h].memory_limit) {
    *bytes = G_cfg->devices[h].total_memory;
    return CUDA_SUCCESS;
  }
  return R.cuDeviceTotalMem_v2 ? R.cuDeviceTotalMem_v2(bytes, dev)
         : R.cuDeviceTotalMem  ? R.cuDeviceTotalMem(bytes, dev)
                               : CUDA_ERROR_NOT_FOUND;
}
VGPU_EXPORT CUresult cuDeviceTotalMem_v2(size_t *b, CUdevice d) { return total_mem(b, d); }
VGPU_EXPORT CUresult cuDeviceTotalMem(size_t *b, CUdevice d) { return total_mem(b, d); }

static CUresult real_meminfo(size_t *fr, size_t *tot) {
  return R.cuMemGetInfo_v2 ? R.cuMemGetInfo_v2(fr, tot) : R.cuMemGetInfo ? R.cuMemGetInfo(fr, tot) : CUDA_ERROR_NOT_FOUND;
}

/* A device without a memory cap is reported as the driver reports it (the reference forwards: cuda_hook.c:1745-1750,
 * nvml_hook.c:57-60) - less what the library instances of this container keep on that device themselves when a core
 * limit made them bring their runtime up there: under the reference that memory would not exist.  Read-only look at the
 * footprint registry (no GPU lock is taken on this path, like the reference). */
static uint64_t own_bytes_on_uncapped(int h) {
  if (h < 0) return 0; /* not one of the container's devices: no runtime is ever brought up there */
  /* a core limit or a ledgered managed allocation (cuMemAllocManaged with GLOBAL attach) brings it up */
  return vgpu_self_registry(h, 0, -1);
}

static CUresult mem_info(size_t *free_out, size_t *total_out) {
  CUdevice dev;
  if (unlikely(!G_cfg)) vgpu_boot();
  CUresult r = R.cuCtxGetDevice ? R.cuCtxGetDevice(&dev) : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) return r;
  int h = vgpu_host_index_of_cuda(dev);
  if (h < 0 || !G_cfg->devices[h].memory_limit) {
    r = real_meminfo(free_out, total_out);
    if (r == CUDA_SUCCESS && free_out && total_out) {
      uint64_t own = own_bytes_on_uncapped(h);
      *free_out = (*free_out + own > *total_out) ? *total_out : *free_out + own;
    }
    return r;
  }
  memview_t mv;
  /* the reference takes the lock and measures first, then (not oversold) asks the driver for
   * its real total to clamp against (:1739-1782); the clamp itself happens in the kernel, so
   * the driver query has to precede the launch */
  size_t rfree = 0, rtotal = 0;
  int real_ok = 0;
  if (!G_cfg->devices[h].memory_oversold) real_ok = real_meminfo(&rfree, &rtotal) == CUDA_SUCCESS;
  memview(&mv, dev, h, NULL, VGPU_Q_CU_INFO, 0, 0, real_ok, rtotal);
  if (mv.failed) {
    r = CUDA_ERROR_NOT_SUPPORTED;
  } else {
    *total_out = mv.res.total;
    *free_out = mv.res.out_free;
    r = CUDA_SUCCESS;
  }
  vgpu_unlock_gpu(mv.lock_fd);
  return r;
}
VGPU_EXPORT CUresult cuMemGetInfo_v2(size_t *f, size_t *t) { return mem_info(f, t); }
VGPU_EXPORT CUresult cuMemGetInfo(size_t *f, size_t *t) { return mem_info(f, t); }

/* ------------------------------------------------------------------ NVML hooks */
static vgpu_dev_rt *rt_for_nvml(int host_index, CUdevice *dev_out) {
  vgpu_dev_rt *rt = vgpu_rt_peek(host_index);
  if (rt) {
    *dev_out = rt->cuda_dev;
    return rt;
  }
  CUdevice dev;
  if (R.cuCtxGetDevice && R.cuCtxGetDevice(&dev) == CUDA_SUCCESS && vgpu_host_index_of_cuda(dev) == host_index) {
    *dev_out = dev;
    return vgpu_rt_get(host_index, dev);
  }
  return NULL; /* no CUDA context in this process: answered on the host (host_nvml_view) */
}

static nvmlReturn_t nvml_view(nvmlDevice_t device, int host_index, vgpu_quota_res_t *out) {
  CUdevice dev = 0;
  vgpu_dev_rt *rt = rt_for_nvml(host_index, &dev);
  if (!rt) {
    /* A context is never created behind the client's back: frameworks query NVML before they
     * fork their CUDA workers precisely to keep the parent CUDA-free. */
    static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    static vgpu_quota_req_t *hq;
    static gfx_scratch_t *hg;
    pthread_mutex_lock(&mu);
    if (!hq) hq = (vgpu_quota_req_t *)calloc(1, sizeof *hq);
    if (!hg) hg = (gfx_scratch_t *)calloc(1, sizeof *hg);
    if (!hq || !hg) {
      pthread_mutex_unlock(&mu);
      return NVML_ERROR_NOT_SUPPORTED;
    }
    int lock_fd = vgpu_lock_gpu(host_index);
    hq->n_graphics = 0;
    if (stage_request(hq, host_index, device)) {
      fetch_graphics(hg, device);
      graphics_install(hq, hg);
    }
    hq->self_bytes = lock_fd >= 0 ? vgpu_self_registry(host_index, 0, 0) : 0;
    host_nvml_view(hq, out);
    vgpu_unlock_gpu(lock_fd);
    pthread_mutex_unlock(&mu);
    return NVML_SUCCESS;
  }
  memview_t mv;
  CUcontext prev = NULL;
  int pushed = 0;
  if (R.cuCtxGetCurrent && R.cuCtxGetCurrent(&prev) == CUDA_SUCCESS && prev != rt->ctx &&
      R.cuCtxPushCurrent_v2 && R.cuCtxPushCurrent_v2(rt->ctx) == CUDA_SUCCESS)
    pushed = 1;
  memview(&mv, dev, host_index, device, VGPU_Q_NVML_INFO, 0, 0, 0, 0);
  if (pushed) R.cuCtxPopCurrent_v2(&prev);
  vgpu_unlock_gpu(mv.lock_fd);
  if (mv.failed || !mv.limited) return NVML_ERROR_NOT_SUPPORTED;
  *out = mv.res;
  return NVML_SUCCESS;
}

VGPU_EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo(nvmlDevice_t device, vnv_memory_t *memory) {
  vgpu_boot();
  int h = vgpu_host_index_of_nvml(device);
  if (h >= 0 && G_cfg->devices[h].memory_limit) {
    vgpu_quota_res_t res;
    nvmlReturn_t r = nvml_view(device, h, &res);
    if (r != NVML_SUCCESS) return r;
    memory->total = res.total;
    memory->used = res.out_used;
    memory->free = res.out_free;
    return NVML_SUCCESS;
  }
  nvmlReturn_t r = R.nvmlDeviceGetMemoryInfo ? R.nvmlDeviceGetMemoryInfo(device, memory) : NVML_ERROR_FUNCTION_NOT_FOUND;
  if (r == NVML_SUCCESS && memory) {
    uint64_t own = own_bytes_on_uncapped(h);
    if (own > memory->used) own = memory->used;
    memory->used -= own;
    memory->free += own;
  }
  return r;
}

VGPU_EXPORT nvmlReturn_t nvmlDeviceGetMemoryInfo_v2(nvmlDevice_t device, vnv_memory_v2_t *memory) {
  vgpu_boot();
  nvmlReturn_t r = R.nvmlDeviceGetMemoryInfo_v2 ? R.nvmlDeviceGetMemoryInfo_v2(device, memory)
                                                : NVML_ERROR_FUNCTION_NOT_FOUND;
  if (r != NVML_SUCCESS) return r;
  int h = vgpu_host_index_of_nvml(device);
  if (h >= 0 && G_cfg->devices[h].memory_limit) {
    vgpu_quota_res_t res;
    r = nvml_view(device, h, &res);
    if (r != NVML_SUCCESS) return r;
    memory->total = res.total; /* version / reserved stay the driver's (nvml_hook.c:89-98) */
    memory->used = res.out_used;
    memory->free = res.out_free;
  } else {
    uint64_t own = own_bytes_on_uncapped(h);
    if (own > memory->used) own = memory->used;
    memory->used -= own;
    memory->free += own;
  }
  return NVML_SUCCESS;
}

VGPU_EXPORT nvmlReturn_t nvmlDeviceSetComputeMode(nvmlDevice_t device, int mode) {
  vgpu_boot();
  int h = vgpu_host_index_of_nvml(device);
  if (h >= 0 && (G_cfg->devices[h].memory_limit || G_cfg->devices[h].core_limit)) return NVML_ERROR_NOT_SUPPORTED;
  return R.nvmlDeviceSetComputeMode ? R.nvmlDeviceSetComputeMode(device, mode) : NVML_ERROR_FUNCTION_NOT_FOUND;
}

VGPU_EXPORT nvmlReturn_t nvmlDeviceGetPersistenceMode(nvmlDevice_t device, int *mode) {
  (void)device;
  *mode = 0; /* NVML_FEATURE_DISABLED, unconditionally (nvml_hook.c:121-126) */
  return NVML_SUCCESS;
}

VGPU_EXPORT nvmlReturn_t nvmlDeviceGetUtilizationRates(nvmlDevice_t device, vnv_utilization_t *u) {
  vgpu_boot(); /* pure forward, like the reference (nvml_originals.c:698-702) */
  return R.nvmlDeviceGetUtilizationRates ? R.nvmlDeviceGetUtilizationRates(device, u)
                                         : NVML_ERROR_FUNCTION_NOT_FOUND;
}
