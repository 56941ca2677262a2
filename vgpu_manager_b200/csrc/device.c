/*
 * device.c - per-GPU device runtime: loads the embedded sm_100a image into the tenant's
 * context through the *real* driver table (never through our own hooks), owns the private
 * streams, the pinned host blocks and the HBM-resident limiter / slab state, and exposes the
 * direct C-ABI entry points declared in include/vgpu_b200.h.
 *
 * No reference counterpart: the reference has no device code.
 */
#include "vgpu_internal.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <time.h>

#include "../../include/vgpu_b200.h"

extern const unsigned char vgpu_kernels_image[]; /* generated from kernels.cu (bin2c) */
extern const unsigned long long vgpu_kernels_image_size;

static vgpu_dev_rt g_rt[VGPU_MAX_DEVICES];
static pthread_mutex_t g_rt_mu = PTHREAD_MUTEX_INITIALIZER;
static volatile unsigned g_rt_epoch;

#define CU_TRY(call, what)                                                             \
  do {                                                                                 \
    CUresult _r = (call);                                                              \
    if (_r != CUDA_SUCCESS) {                                                          \
      VLOG(VL_ERROR, "device runtime: %s failed: %d (%s)", what, _r, vgpu_cu_err(_r)); \
      goto fail;                                                                       \
    }                                                                                  \
  } while (0)

static int pinned_block(size_t bytes, void **host, CUdeviceptr *dev) {
  if (!R.cuMemHostAlloc || !R.cuMemHostGetDevicePointer_v2) return -1;
  if (R.cuMemHostAlloc(host, bytes, VCU_MEMHOSTALLOC_PORTABLE | VCU_MEMHOSTALLOC_DEVICEMAP)) return -1;
  memset(*host, 0, bytes);
  if (R.cuMemHostGetDevicePointer_v2(dev, *host, 0)) return -1;
  return 0;
}

CUresult vgpu_rt_launch(vgpu_dev_rt *rt, CUfunction f, unsigned grid, unsigned block, unsigned smem,
                        CUstream s, void **params) {
  (void)rt;
  return VGPU_CAPCHK(R.cuLaunchKernel(f, grid, 1, 1, block, 1, 1, smem, s, params, NULL));
}

/* bytes NVML attributes to this process on `nvdev` (calibration of our own footprint only) */
static uint64_t own_process_bytes(nvmlDevice_t nvdev) {
  if (!nvdev || !R.nvmlDeviceGetComputeRunningProcesses) return 0;
  static __thread vgpu_proc_t procs[VGPU_MAX_PIDS];
  unsigned int n = VGPU_MAX_PIDS;
  if (R.nvmlDeviceGetComputeRunningProcesses(nvdev, &n, procs) != NVML_SUCCESS) return 0;
  uint32_t me = (uint32_t)getpid();
  for (unsigned int i = 0; i < n; i++)
    if (procs[i].pid == me) return procs[i].used_bytes;
  /* pid namespace: our host pid is unknown; with a single process on the device it is us */
  return n == 1 ? procs[0].used_bytes : 0;
}

static int spin_seq(volatile uint32_t *word, uint32_t want, CUstream s) {
  /* the kernel publishes its sequence number with a system-scope release; polling the pinned
   * word is ~2x cheaper than cuStreamSynchronize for a ~5 us kernel */
  for (int i = 0; i < 2000000; i++) {
    if (*word == want) return 0;
    if ((i & 1023) == 1023) {
      int prev = vgpu_capture_relax();
      CUresult q = R.cuStreamQuery ? VGPU_CAPCHK(R.cuStreamQuery(s)) : CUDA_SUCCESS;
      vgpu_capture_restore(prev);
      if (q != CUDA_SUCCESS && q != CUDA_ERROR_NOT_READY) return -1;
      sched_yield();
    }
    __builtin_ia32_pause();
  }
  int prev = vgpu_capture_relax();
  CUresult sr = VGPU_CAPCHK(R.cuStreamSynchronize(s));
  vgpu_capture_restore(prev);
  if (sr != CUDA_SUCCESS) return -1;
  return *word == want ? 0 : -1;
}

/* How the controller's utilisation reading is formed.  Default `nvml`: the reference's own
 * signal - per-process NVML samples (or sm_util.config), published once per control period and
 * folded on the device by vgpu_refill_kernel.  `queue` / `sm` / `max` are the driver-free
 * on-device signals (stream queue-busy, per-SM issue-slot probe). */
static void read_util_tunables(vgpu_lim_host_t *H) {
  const char *src = vgpu_tunable("VGPU_B200_UTIL_SOURCE");
  H->util_source = !src ? VGPU_SRC_NVML : !strcmp(src, "queue") ? VGPU_SRC_QUEUE : !strcmp(src, "sm") ? VGPU_SRC_SM
                   : !strcmp(src, "max") ? VGPU_SRC_MAX : VGPU_SRC_NVML;
  const char *win = vgpu_tunable("VGPU_B200_UTIL_WINDOW_PERIODS"); /* control periods (~80 ms) per utilisation reading */
  H->util_window = (win && atoi(win) >= 1) ? (uint32_t)atoi(win) : 4u;
  const char *um = vgpu_tunable("VGPU_B200_UTIL_MODE"); /* block (default, NVML-like) | average */
  H->util_mode = (um && !strcmp(um, "average")) ? 0u : 1u;
}

static void write_limiter_config(vgpu_dev_rt *rt, vgpu_lim_dev_t *init) {
  const vgpu_cfg_dev_t *c = (rt->host_index >= 0 && G_cfg) ? &G_cfg->devices[rt->host_index] : NULL;
  memset(init, 0, sizeof *init);
  init->sm_num = rt->sm_num;
  init->max_thread_per_sm = rt->max_thread_per_sm;
  init->total_cores = rt->total_cores;
  init->pre_sys_process_num = 1;
  if (c) {
    init->hard_core = c->hard_core;
    init->soft_core = c->soft_core;
    init->core_limit = c->core_limit;
    init->hard_limit = c->hard_limit;
    init->up_limit = c->hard_core;
  }
  for (unsigned i = 0; i < VGPU_MAX_SMS * 4; i++) init->probe_idle[i] = 0xffffffffu;
}

static vgpu_dev_rt *bring_up(vgpu_dev_rt *rt, int slot, int host_index, CUdevice dev) {
  CUcontext ctx = NULL;
  int retained = 0;
  if (!R.cuCtxGetCurrent || R.cuCtxGetCurrent(&ctx) != CUDA_SUCCESS || !ctx) {
    /* no current context on this thread (e.g. cuMemCreate addressed the device through
     * prop->location): fall back to the device's primary context */
    if (!R.cuDevicePrimaryCtxRetain || R.cuDevicePrimaryCtxRetain(&ctx, dev) != CUDA_SUCCESS || !ctx ||
        R.cuCtxPushCurrent_v2(ctx) != CUDA_SUCCESS)
      return NULL;
    retained = 1;
  }
  if (!R.cuModuleLoadData || !R.cuLaunchKernel || !R.cuMemAlloc_v2) {
    VLOG(VL_ERROR, "device runtime: driver lacks module/launch entry points");
    if (retained) { CUcontext dummy; R.cuCtxPopCurrent_v2(&dummy); }
    rt->ready = -1;
    rt->retry_at = 0; /* deterministic: never retried */
    return NULL;
  }
  unsigned fails = rt->fails;
  memset(rt, 0, sizeof *rt);
  rt->fails = fails;
  pthread_mutex_init(&rt->q_mu, NULL);
  rt->host_index = host_index;
  rt->cuda_dev = dev;
  rt->ctx = ctx;
  {
    unsigned int fl = 0;
    int active = 0;
    CUcontext prim = NULL;
    rt->ctx_is_primary = retained;
    if (!retained && R.cuDevicePrimaryCtxGetState && R.cuDevicePrimaryCtxGetState(dev, &fl, &active) == CUDA_SUCCESS && active &&
        R.cuDevicePrimaryCtxRetain && R.cuDevicePrimaryCtxRetain(&prim, dev) == CUDA_SUCCESS) {
      rt->ctx_is_primary = (prim == ctx);
      if (R.cuDevicePrimaryCtxRelease_v2) R.cuDevicePrimaryCtxRelease_v2(dev);
      else if (R.cuDevicePrimaryCtxRelease) R.cuDevicePrimaryCtxRelease(dev);
    }
  }

  nvmlDevice_t nvdev = vgpu_nvml_handle_of_host(host_index);
  /* The footprint measurement compares NVML's figure for THIS pid before and after: other
   * processes cannot disturb it, so the per-GPU lock is only taken to publish the result.  (Holding it
   * across module load + warm-ups serialised the bring-up of tenants that start together by
   * 0.1-0.3 s each; the tenant that came up first then ramped its share alone and kept a
   * multi-million-token head start for the rest of the run - the controller's increments are the
   * same for everybody once they see the same reading, so start-up offsets never decay.  Measured as
   * max/min = 4-8 between four 25 % tenants in one run out of three.) */
  int lock_fd = -1;
  uint64_t before = own_process_bytes(nvdev);

  R.cuDeviceGetAttribute(&rt->sm_num, VCU_ATTR_SM_COUNT, dev);
  R.cuDeviceGetAttribute(&rt->max_thread_per_sm, VCU_ATTR_MAX_THREADS_PER_SM, dev);
  rt->total_cores = (int64_t)rt->max_thread_per_sm * (int64_t)rt->sm_num * 32; /* cuda_hook.c:534 */
  int m64 = 0;
  R.cuDeviceGetAttribute(&m64, 122 /* CAN_USE_64_BIT_STREAM_MEM_OPS */, dev);
  rt->memops64 = m64 && R.cuStreamWaitValue64_v2 && R.cuStreamWriteValue64_v2;

  CU_TRY(R.cuModuleLoadData(&rt->mod, vgpu_kernels_image), "cuModuleLoadData(sm_100a image)");
  CU_TRY(R.cuModuleGetFunction(&rt->k_clear, rt->mod, VGPU_K_CLEAR), VGPU_K_CLEAR);
  CU_TRY(R.cuModuleGetFunction(&rt->k_spill, rt->mod, VGPU_K_SPILL), VGPU_K_SPILL);
  CU_TRY(R.cuModuleGetFunction(&rt->k_copy_generic, rt->mod, "vgpu_copy_generic_kernel"), "generic copy");
  CU_TRY(R.cuModuleGetFunction(&rt->k_quota, rt->mod, VGPU_K_QUOTA), VGPU_K_QUOTA);
  CU_TRY(R.cuModuleGetFunction(&rt->k_slab_insert, rt->mod, VGPU_K_SLAB_INSERT), VGPU_K_SLAB_INSERT);
  CU_TRY(R.cuModuleGetFunction(&rt->k_slab_remove, rt->mod, VGPU_K_SLAB_REMOVE), VGPU_K_SLAB_REMOVE);
  CU_TRY(R.cuModuleGetFunction(&rt->k_controller, rt->mod, VGPU_K_CONTROLLER), VGPU_K_CONTROLLER);
  CU_TRY(R.cuModuleGetFunction(&rt->k_sampler, rt->mod, VGPU_K_SAMPLER), VGPU_K_SAMPLER);
  CU_TRY(R.cuModuleGetFunction(&rt->k_gate, rt->mod, VGPU_K_GATE), VGPU_K_GATE);
  CU_TRY(R.cuModuleGetFunction(&rt->k_governor, rt->mod, VGPU_K_GOVERNOR), VGPU_K_GOVERNOR);
  CU_TRY(R.cuModuleGetFunction(&rt->k_refill, rt->mod, VGPU_K_REFILL), VGPU_K_REFILL);
  CU_TRY(R.cuModuleGetFunction(&rt->k_vslab, rt->mod, VGPU_K_VSLAB), VGPU_K_VSLAB);
  {
    /* spill-copy geometry: defaults from kernel_abi.h, overridable for tuning sweeps */
    const char *e;
    rt->spill_chunk = VGPU_SPILL_CHUNK;
    rt->spill_stages = VGPU_SPILL_STAGES;
    rt->spill_ctas_per_sm = VGPU_SPILL_CTAS_PER_SM;
    if ((e = getenv("VGPU_B200_SPILL_CHUNK")) && atoi(e) >= 1024) rt->spill_chunk = (uint32_t)atoi(e) & ~15u;
    if ((e = getenv("VGPU_B200_SPILL_STAGES")) && atoi(e) >= 2 && atoi(e) <= 16) rt->spill_stages = (uint32_t)atoi(e);
    if ((e = getenv("VGPU_B200_SPILL_CTAS_PER_SM")) && atoi(e) >= 1) rt->spill_ctas_per_sm = (uint32_t)atoi(e);
    if ((size_t)rt->spill_chunk * rt->spill_stages > 200 * 1024) rt->spill_stages = 200 * 1024 / rt->spill_chunk;
  }
  if (R.cuFuncSetAttribute) /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8 */
    CU_TRY(R.cuFuncSetAttribute(rt->k_spill, 8, (int)(rt->spill_chunk * rt->spill_stages)), "spill smem opt-in");

  /* One private stream by default: every extra stream of a context makes the driver's own
   * device-wide operations (cuMemFree, cuMemAlloc) dearer for the tenant.  The resident governor
   * and the per-SM probe get streams of their own only when those modes are asked for. */
  int lo = 0, hi = 0;
  if (R.cuCtxGetStreamPriorityRange) R.cuCtxGetStreamPriorityRange(&lo, &hi);
  {
    const char *gov = vgpu_tunable("VGPU_B200_GOVERNOR"), *src = vgpu_tunable("VGPU_B200_UTIL_SOURCE");
    /* (the on-device signals keep sampler windows of 0.5 ms resident: a quota evaluation must not queue behind one) */
    int separate = (gov && atoi(gov)) || (src && (!strcmp(src, "queue") || !strcmp(src, "sm") || !strcmp(src, "max")));
    if (R.cuStreamCreateWithPriority) CU_TRY(R.cuStreamCreateWithPriority(&rt->q_stream, VCU_STREAM_NON_BLOCKING, hi), "library stream");
    else CU_TRY(R.cuStreamCreate(&rt->q_stream, VCU_STREAM_NON_BLOCKING), "library stream");
    rt->s_stream = rt->p_stream = rt->q_stream;
    if (separate) {
      if (R.cuStreamCreateWithPriority) {
        CU_TRY(R.cuStreamCreateWithPriority(&rt->s_stream, VCU_STREAM_NON_BLOCKING, hi), "governor stream");
        CU_TRY(R.cuStreamCreateWithPriority(&rt->p_stream, VCU_STREAM_NON_BLOCKING, hi), "probe stream");
      } else {
        CU_TRY(R.cuStreamCreate(&rt->s_stream, VCU_STREAM_NON_BLOCKING), "governor stream");
        CU_TRY(R.cuStreamCreate(&rt->p_stream, VCU_STREAM_NON_BLOCKING), "probe stream");
      }
    }
  }

  if (pinned_block(sizeof(vgpu_quota_req_t), (void **)&rt->q_req, &rt->q_req_d) ||
      pinned_block(sizeof(vgpu_quota_res_t), (void **)&rt->q_res, &rt->q_res_d) ||
      pinned_block(sizeof(vgpu_slab_res_t), (void **)&rt->slab_res, &rt->slab_res_d) ||
      pinned_block(sizeof(vgpu_util_req_t), (void **)&rt->u_req, &rt->u_req_d) ||
      pinned_block(sizeof(vgpu_vslab_res_t), (void **)&rt->vs_res, &rt->vs_res_d) ||
      pinned_block(sizeof(vgpu_lim_host_t), (void **)&rt->lim_h, &rt->lim_h_d)) {
    VLOG(VL_ERROR, "device runtime: pinned host blocks unavailable");
    goto fail;
  }
  rt->lim_h->ext_user_override = -1;
  rt->lim_h->ext_sys_process_num = 1;
  read_util_tunables(rt->lim_h);

  /* One HBM allocation: limiter state followed by the UVA slab, rounded up to a whole 2 MiB
   * allocation granule.  A sub-granule request would be carved from the driver's small-block
   * pool, whose free tail the tenant's own small allocations would then share - their first
   * 2 MiB granule would stop showing up in NVML and the reported usage would differ from a
   * reference deployment by one granule (seen with the reference's test_alloc). */
  size_t lim_bytes = (sizeof(vgpu_lim_dev_t) + 255) & ~(size_t)255;
  size_t slab_bytes = sizeof(vgpu_slab_slot_t) * VGPU_SLAB_SLOTS;
  size_t hbm_bytes = lim_bytes + slab_bytes + sizeof(vgpu_vslab_slot_t) * VGPU_VSLAB_SLOTS;
  hbm_bytes = (hbm_bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  CU_TRY(R.cuMemAlloc_v2(&rt->lim_d, hbm_bytes), "HBM state");
  rt->slab_d = rt->lim_d + lim_bytes;
  rt->vslab_d = rt->slab_d + slab_bytes;
  CU_TRY(R.cuMemsetD8_v2(rt->lim_d, 0, hbm_bytes), "HBM state clear");
  {
    vgpu_lim_dev_t *init = (vgpu_lim_dev_t *)malloc(sizeof *init);
    if (!init) goto fail;
    write_limiter_config(rt, init);
    CUresult r = R.cuMemcpyHtoD_v2(rt->lim_d, init, sizeof *init);
    free(init);
    CU_TRY(r, "HBM state init");
  }

  /* force lazy module loading to materialise every kernel now, so that the footprint we
   * measure below is final: null-work launches */
  {
    unsigned long long zero = 0;
    CUdeviceptr nullp = rt->slab_d;
    void *p_clear[] = {&nullp, &zero};
    void *p_copy[] = {&nullp, &nullp, &zero, &rt->spill_chunk, &rt->spill_stages};
    CU_TRY(vgpu_rt_launch(rt, rt->k_clear, 1, 256, 0, rt->q_stream, p_clear), "warm clear");
    CU_TRY(vgpu_rt_launch(rt, rt->k_spill, 1, 32, rt->spill_chunk * rt->spill_stages, rt->q_stream, p_copy), "warm spill");
    CU_TRY(vgpu_rt_launch(rt, rt->k_copy_generic, 1, 256, 0, rt->q_stream, p_copy), "warm copy");
    rt->q_req->seq = ++rt->seq;
    uint32_t plain = 0;
    void *p_quota[] = {&rt->q_req_d, &rt->q_res_d, &plain};
    CU_TRY(vgpu_rt_launch(rt, rt->k_quota, 1, 1024, 0, rt->q_stream, p_quota), "warm quota");
    uint32_t sq = ++rt->seq;
    unsigned long long key = 2, bytes = 0;
    void *p_ins[] = {&rt->slab_d, &key, &bytes, &rt->slab_res_d, &sq};
    void *p_rem[] = {&rt->slab_d, &key, &rt->slab_res_d, &sq};
    CU_TRY(vgpu_rt_launch(rt, rt->k_slab_insert, 1, 32, 0, rt->q_stream, p_ins), "warm slab insert");
    CU_TRY(vgpu_rt_launch(rt, rt->k_slab_remove, 1, 32, 0, rt->q_stream, p_rem), "warm slab remove");
    vgpu_ctrl_in_t in = {0, 0, 0, 1};
    void *p_ctl[] = {&rt->lim_d, &rt->lim_h_d, &in};
    CU_TRY(vgpu_rt_launch(rt, rt->k_controller, 1, 32, 0, rt->q_stream, p_ctl), "warm controller");
    vgpu_vslab_req_t vq = {VGPU_VSLAB_TAKE, 0, 0, 0, 0, 0, 2, 0, 0, 0, 0};
    void *p_vs[] = {&rt->vslab_d, &vq, &rt->vs_res_d, &sq};
    CU_TRY(vgpu_rt_launch(rt, rt->k_vslab, 1, 1024, 0, rt->q_stream, p_vs), "warm slab table");
    void *p_ref[] = {&rt->lim_d, &rt->lim_h_d, &rt->u_req_d};
    CU_TRY(vgpu_rt_launch(rt, rt->k_refill, 1, 32, 0, rt->q_stream, p_ref), "warm refill");
    uint32_t w = 0, iv = 1, per = VGPU_SAMPLER_PROBE_ONLY, ep = 0;
    uint32_t none = 0;
    void *p_smp[] = {&rt->lim_d, &rt->lim_h_d, &w, &iv, &per, &ep, &none};
    CU_TRY(vgpu_rt_launch(rt, rt->k_sampler, 1, 128, 0, rt->q_stream, p_smp), "warm sampler");
    rt->lim_h->quit = 1;
    rt->lim_h->ctl_state = 1;
    uint32_t g_iv = 10, g_per = 80000, g_idle = 1;
    void *p_gov[] = {&rt->lim_d, &rt->lim_h_d, &g_iv, &g_per, &g_idle};
    CU_TRY(vgpu_rt_launch(rt, rt->k_governor, 1, 32, 0, rt->q_stream, p_gov), "warm governor");
    long long tk = 0;
    uint32_t to = 1;
    void *p_gate[] = {&rt->lim_d, &tk, &to};
    CU_TRY(vgpu_rt_launch(rt, rt->k_gate, 1, 1, 0, rt->q_stream, p_gate), "warm gate");
    CU_TRY(R.cuStreamSynchronize(rt->q_stream), "warm-up sync");
    for (int i = 0; i < 10000 && rt->lim_h->ctl_state != 0; i++) { /* governor's last store is to host memory */
      struct timespec nap = {0, 100000};
      nanosleep(&nap, NULL);
    }
    /* the warm-up controller/sampler steps touched the state: re-initialise it */
    CU_TRY(R.cuMemsetD8_v2(rt->lim_d, 0, lim_bytes), "HBM state re-clear");
    vgpu_lim_dev_t *init = (vgpu_lim_dev_t *)malloc(sizeof *init);
    if (!init) goto fail;
    write_limiter_config(rt, init);
    CUresult r = R.cuMemcpyHtoD_v2(rt->lim_d, init, sizeof *init);
    free(init);
    CU_TRY(r, "HBM state init");
    memset((void *)rt->lim_h, 0, offsetof(vgpu_lim_host_t, user_current));
    rt->lim_h->gov_left_busy = 0;
    rt->lim_h->ext_user_override = -1;
    rt->lim_h->ext_sys_process_num = 1;
    read_util_tunables(rt->lim_h);
  }

  uint64_t after = own_process_bytes(nvdev);
  rt->self_bytes = after > before ? after - before : 0;
  lock_fd = host_index >= 0 ? vgpu_lock_gpu(host_index) : -1;
  if (host_index >= 0 && lock_fd >= 0) vgpu_self_registry(host_index, rt->self_bytes, 1);
  vgpu_unlock_gpu(lock_fd);
  rt->fails = 0;
  rt->retry_at = 0;
  {
    const char *e = getenv("VGPU_B200_QUOTA_ARMED"); /* 0: launch the quota kernel after the NVML queries (round-1 behaviour) */
    rt->quota_armed = !(e && *e == '0');
  }
  __sync_synchronize();
  rt->ready = 1;
  if (retained) { CUcontext dummy; R.cuCtxPopCurrent_v2(&dummy); }
  VLOG(VL_INFO, "device runtime up: slot %d host %d sms %d own footprint %" PRIu64 " bytes memops64 %d",
       slot, host_index, rt->sm_num, rt->self_bytes, rt->memops64);
  return rt;
fail:
  vgpu_unlock_gpu(lock_fd);
  /* give back whatever was created: a transient cause (the tenant filled the GPU, a capture
   * conflict) must not leak a module, three streams and the pinned blocks on every retry */
  if (rt->lim_d && R.cuMemFree_v2) R.cuMemFree_v2(rt->lim_d);
  if (R.cuMemFreeHost) {
    if (rt->q_req) R.cuMemFreeHost(rt->q_req);
    if (rt->q_res) R.cuMemFreeHost(rt->q_res);
    if (rt->slab_res) R.cuMemFreeHost(rt->slab_res);
    if (rt->u_req) R.cuMemFreeHost(rt->u_req);
    if (rt->vs_res) R.cuMemFreeHost(rt->vs_res);
    if (rt->lim_h) R.cuMemFreeHost((void *)rt->lim_h);
  }
  if (R.cuStreamDestroy_v2) {
    if (rt->s_stream && rt->s_stream != rt->q_stream) R.cuStreamDestroy_v2(rt->s_stream);
    if (rt->p_stream && rt->p_stream != rt->q_stream) R.cuStreamDestroy_v2(rt->p_stream);
    if (rt->q_stream) R.cuStreamDestroy_v2(rt->q_stream);
  }
  if (rt->mod && R.cuModuleUnload) R.cuModuleUnload(rt->mod);
  if (retained) {
    CUcontext dummy;
    R.cuCtxPopCurrent_v2(&dummy);
    if (R.cuDevicePrimaryCtxRelease_v2) R.cuDevicePrimaryCtxRelease_v2(dev);
  }
  {
    unsigned f = rt->fails + 1;
    pthread_mutex_destroy(&rt->q_mu);
    memset(rt, 0, sizeof *rt);
    rt->fails = f;
    /* back-off 1 s, 2 s, 4 s ... 32 s: the next hooked call after that tries again */
    struct timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
    rt->retry_at = (uint64_t)now.tv_sec + (1ull << (f > 6 ? 5 : f - 1));
    rt->ready = -1;
  }
  VLOG(VL_ERROR, "device runtime bring-up failed on cuda device %d (attempt %u): the sm_100a enforcement "
                 "kernels are unavailable until the retry (no CPU fallback exists)", dev, rt->fails);
  return NULL;
}

vgpu_dev_rt *vgpu_rt_get(int host_index, CUdevice dev) {
  int slot = host_index >= 0 ? host_index : (dev >= 0 && dev < VGPU_MAX_DEVICES ? dev : 0);
  unsigned me = vgpu_fork_epoch + 1;
  vgpu_dev_rt *rt = &g_rt[slot];
  if (likely(g_rt_epoch == me && rt->ready == 1)) return rt;
  pthread_mutex_lock(&g_rt_mu);
  if (g_rt_epoch != me) { /* fork: device state does not survive, start over */
    memset(g_rt, 0, sizeof g_rt);
    g_rt_epoch = me;
  }
  vgpu_dev_rt *out = NULL;
  if (rt->ready == -1 && rt->retry_at) { /* a transient failure is retried after its back-off */
    struct timespec now;
    clock_gettime(CLOCK_MONOTONIC, &now);
    if ((uint64_t)now.tv_sec >= rt->retry_at) rt->ready = 0;
  }
  if (rt->ready == 1) out = rt;
  else if (rt->ready == 0) {
    /* module load, allocations and the warm-up synchronise are "unsafe" calls for CUDA's capture
     * rules: another tenant thread may be capturing in global mode right now */
    int prev = vgpu_capture_relax();
    out = bring_up(rt, slot, host_index, dev);
    vgpu_capture_restore(prev);
  }
  pthread_mutex_unlock(&g_rt_mu);
  return out;
}

/* ------------------------------------------------------------------ UVA records that outlive their context
 * The reference keeps its UVA allocation nodes in a host list (loader.c:1824-1907) which a device reset does not
 * touch: a node whose memory went with the context stays listed, and when a later allocation that happens to get
 * the same address is freed, free_gpu_virt_memory finds the stale node and shrinks the ledger by its size.  The
 * records here live in HBM and would vanish with the context, so they are read back once, while the context is
 * being torn down (one 256 KiB copy on the library's own stream), and kept as the "stale" list a free falls back
 * to after the live tables.  Newest first, like list_add: live records are always newer than stale ones, a later
 * reset's records newer than an earlier one's. */
static vgpu_slab_slot_t *g_stale;
static uint32_t g_stale_n, g_stale_cap;
static pthread_mutex_t g_stale_mu = PTHREAD_MUTEX_INITIALIZER;

int vgpu_stale_uva_remove(CUdeviceptr dptr, uint64_t *bytes) {
  if (!g_stale_n) return 1;
  int rc = 1;
  pthread_mutex_lock(&g_stale_mu);
  for (uint32_t i = g_stale_n; i-- > 0;)
    if (g_stale[i].dptr == (uint64_t)dptr) {
      *bytes = g_stale[i].bytes;
      memmove(&g_stale[i], &g_stale[i + 1], (size_t)(g_stale_n - i - 1) * sizeof g_stale[0]);
      g_stale_n--;
      rc = 0;
      break;
    }
  pthread_mutex_unlock(&g_stale_mu);
  return rc;
}

static void stale_keep(vgpu_slab_slot_t *recs, uint32_t n) {
  pthread_mutex_lock(&g_stale_mu);
  if (g_stale_n + n > g_stale_cap) {
    uint32_t cap = (g_stale_n + n + 63u) & ~63u;
    vgpu_slab_slot_t *p = (vgpu_slab_slot_t *)realloc(g_stale, (size_t)cap * sizeof *p);
    if (p) { g_stale = p; g_stale_cap = cap; }
  }
  if (g_stale_n + n <= g_stale_cap) {
    memcpy(&g_stale[g_stale_n], recs, (size_t)n * sizeof recs[0]);
    g_stale_n += n;
  }
  pthread_mutex_unlock(&g_stale_mu);
}

static int ctx_enter(vgpu_dev_rt *rt);
static void ctx_leave(int pushed);

/* caller holds g_rt_mu; the runtime is still whole */
static void uva_snapshot(vgpu_dev_rt *rt) {
  rt->uva_snap = NULL;
  rt->uva_snap_n = 0;
  if (rt->uva_live <= 0 || !R.cuMemcpyDtoHAsync_v2 || !R.cuStreamSynchronize) return;
  vgpu_slab_slot_t *tab = (vgpu_slab_slot_t *)malloc(sizeof(vgpu_slab_slot_t) * VGPU_SLAB_SLOTS);
  if (!tab) return;
  pthread_mutex_lock(&rt->q_mu);
  int pushed = ctx_enter(rt);
  /* the library's own non-blocking stream: not ordered behind anything of the tenant's */
  CUresult r = R.cuMemcpyDtoHAsync_v2(tab, rt->slab_d, sizeof(vgpu_slab_slot_t) * VGPU_SLAB_SLOTS, rt->q_stream);
  if (r == CUDA_SUCCESS) r = R.cuStreamSynchronize(rt->q_stream);
  ctx_leave(pushed);
  pthread_mutex_unlock(&rt->q_mu);
  if (r != CUDA_SUCCESS) {
    VLOG(VL_WARNING, "could not read the UVA records back before the context goes (%d: %s)", r, vgpu_cu_err(r));
    free(tab);
    return;
  }
  uint32_t n = 0;
  for (uint32_t i = 0; i < VGPU_SLAB_SLOTS; i++)
    if (tab[i].dptr > 1) tab[n++] = tab[i]; /* 0 free, 1 tombstone */
  rt->uva_snap = tab;
  rt->uva_snap_n = n;
}

unsigned vgpu_rt_context_before(CUcontext ctx, CUdevice dev, int primary) {
  unsigned mask = 0;
  if (g_rt_epoch != vgpu_fork_epoch + 1) return 0;
  pthread_mutex_lock(&g_rt_mu);
  for (int slot = 0; slot < VGPU_MAX_DEVICES; slot++) {
    vgpu_dev_rt *rt = &g_rt[slot];
    if (rt->ready != 1) continue;
    int hit = primary ? (rt->cuda_dev == dev && rt->ctx_is_primary) : (rt->ctx == ctx);
    if (!hit) continue;
    if (rt->host_index >= 0) vgpu_limiter_detach(rt->host_index);
    uva_snapshot(rt);
    rt->ready = 2; /* parked: neither usable nor up for a new bring-up until _after() decides */
    mask |= 1u << slot;
  }
  pthread_mutex_unlock(&g_rt_mu);
  return mask;
}

void vgpu_rt_context_after(unsigned mask, int still_alive) {
  if (!mask) return;
  int gone[VGPU_MAX_DEVICES], n_gone = 0;
  pthread_mutex_lock(&g_rt_mu);
  for (int slot = 0; slot < VGPU_MAX_DEVICES; slot++) {
    if (!(mask & (1u << slot))) continue;
    vgpu_dev_rt *rt = &g_rt[slot];
    if (rt->ready != 2) continue;
    if (still_alive) {
      free(rt->uva_snap); /* the table itself is still there */
      rt->uva_snap = NULL;
      rt->ready = 1;
      if (rt->host_index >= 0) vgpu_limiter_attach(rt->host_index, 0);
    } else {
      if (rt->uva_snap_n) stale_keep(rt->uva_snap, rt->uva_snap_n);
      free(rt->uva_snap);
      /* everything the runtime owned went with the context; the next hooked call in a new
       * context brings a fresh one up (token bucket and slab start empty, like a new process) */
      VLOG(VL_INFO, "context of runtime slot %d is gone; device state will be rebuilt on next use", slot);
      if (rt->host_index >= 0) vgpu_limiter_attach(rt->host_index, 1);
      if (rt->host_index >= 0 && rt->self_bytes) gone[n_gone++] = rt->host_index;
      vgpu_slab_forget(rt);
      pthread_mutex_destroy(&rt->q_mu);
      memset(rt, 0, sizeof *rt);
    }
  }
  pthread_mutex_unlock(&g_rt_mu);
  /* the device memory went with the context: take this process's share out of the container's footprint registry
   * (outside g_rt_mu - the allocation path takes the GPU lock first and the runtime table second) */
  for (int i = 0; i < n_gone; i++) {
    int fd = vgpu_lock_gpu(gone[i]);
    if (fd >= 0) {
      vgpu_self_registry(gone[i], 0, 1);
      vgpu_unlock_gpu(fd);
    }
  }
}

vgpu_dev_rt *vgpu_rt_peek(int host_index) {
  if (host_index < 0 || host_index >= VGPU_MAX_DEVICES) return NULL;
  vgpu_dev_rt *rt = &g_rt[host_index];
  return (g_rt_epoch == vgpu_fork_epoch + 1 && rt->ready == 1) ? rt : NULL;
}

/* ------------------------------------------------------------------ kernel drivers used by the hooks */
/* The calling thread may have no current context (cuMemCreate from a worker thread addresses
 * the device through prop->location) or a different one: run our kernels in the runtime's. */
static int ctx_enter(vgpu_dev_rt *rt) {
  CUcontext cur = NULL;
  if (R.cuCtxGetCurrent && R.cuCtxGetCurrent(&cur) == CUDA_SUCCESS && cur == rt->ctx) return 0;
  return (R.cuCtxPushCurrent_v2 && R.cuCtxPushCurrent_v2(rt->ctx) == CUDA_SUCCESS) ? 1 : 0;
}
static void ctx_leave(int pushed) {
  CUcontext dummy;
  if (pushed) R.cuCtxPopCurrent_v2(&dummy);
}

static unsigned quota_block(uint32_t longest) {
  if (longest > VGPU_MAX_PIDS) longest = VGPU_MAX_PIDS;
  return longest ? (longest + 31u) & ~31u : 32u;
}

int vgpu_rt_quota(vgpu_dev_rt *rt, vgpu_quota_res_t *out) {
  /* caller filled rt->q_req (except seq) and holds rt->q_mu */
  uint32_t seq = ++rt->seq;
  rt->q_req->seq = seq;
  if (!rt->q_req_self_set) rt->q_req->self_bytes = rt->self_bytes;
  rt->q_req_self_set = 0;
  __sync_synchronize();
  uint32_t plain = 0;
  void *params[] = {&rt->q_req_d, &rt->q_res_d, &plain};
  int pushed = ctx_enter(rt);
  /* one thread per record of the longest list, whole warps (see the kernel) */
  uint32_t longest = rt->q_req->n_compute > rt->q_req->n_graphics ? rt->q_req->n_compute : rt->q_req->n_graphics;
  if (rt->q_req->n_vmem > longest) longest = rt->q_req->n_vmem;
  rt->q_longest = longest;
  CUresult r = vgpu_rt_launch(rt, rt->k_quota, 1, quota_block(longest), 0, rt->q_stream, params);
  if (r != CUDA_SUCCESS) {
    ctx_leave(pushed);
    VLOG(VL_ERROR, "quota kernel launch failed: %d (%s)", r, vgpu_cu_err(r));
    return -1;
  }
  int stuck = spin_seq(&rt->q_res->seq_done, seq, rt->q_stream);
  ctx_leave(pushed);
  if (stuck) {
    VLOG(VL_ERROR, "quota kernel did not complete");
    return -1;
  }
  *out = *rt->q_res;
  vgpu_metric_add(rt->host_index, VM_QUOTA_KERNELS, 1);
  return 0;
}

/* Armed evaluation, used by the allocation hooks: the kernel is launched first and waits on the
 * device for the request block to be published under `seq`, so its launch latency overlaps the
 * NVML queries the host makes in between.  arm -> (host stages the request) -> publish ->
 * (host may do other work) -> collect.  Caller holds rt->q_mu.  Returns the sequence number, 0 if
 * the launch failed (the caller then uses vgpu_rt_quota). */
uint32_t vgpu_rt_quota_arm(vgpu_dev_rt *rt) {
  uint32_t seq = ++rt->seq;
  if (!seq) seq = ++rt->seq;
  void *params[] = {&rt->q_req_d, &rt->q_res_d, &seq};
  int pushed = ctx_enter(rt);
  /* the lists are not known yet: size the CTA for what the previous call saw plus some slack;
   * a longer list makes the kernel answer VGPU_PATH_RETRY */
  CUresult r = vgpu_rt_launch(rt, rt->k_quota, 1, quota_block(rt->q_longest + 8), 0, rt->q_stream, params);
  ctx_leave(pushed);
  if (r != CUDA_SUCCESS) return 0;
  vgpu_metric_add(rt->host_index, VM_QUOTA_KERNELS, 1);
  return seq;
}

void vgpu_rt_quota_publish(vgpu_dev_rt *rt, uint32_t seq) {
  __sync_synchronize(); /* request first, then the sequence number the kernel waits for */
  *(volatile uint32_t *)&rt->q_req->seq = seq;
}

/* 0 = result in *out; 1 = the kernel asked for a plain re-evaluation; -1 = failure */
int vgpu_rt_quota_collect(vgpu_dev_rt *rt, uint32_t seq, vgpu_quota_res_t *out) {
  int pushed = ctx_enter(rt);
  int stuck = spin_seq(&rt->q_res->seq_done, seq, rt->q_stream);
  ctx_leave(pushed);
  if (stuck) return -1;
  *out = *rt->q_res;
  uint32_t longest = rt->q_req->n_compute > rt->q_req->n_graphics ? rt->q_req->n_compute : rt->q_req->n_graphics;
  if (rt->q_req->n_vmem > longest) longest = rt->q_req->n_vmem;
  rt->q_longest = longest;
  return out->path == VGPU_PATH_RETRY ? 1 : 0;
}

int vgpu_rt_slab_insert(vgpu_dev_rt *rt, CUdeviceptr dptr, uint64_t bytes) {
  pthread_mutex_lock(&rt->q_mu);
  uint32_t seq = ++rt->seq;
  unsigned long long k = dptr, b = bytes;
  void *params[] = {&rt->slab_d, &k, &b, &rt->slab_res_d, &seq};
  int rc = -1;
  int pushed = ctx_enter(rt);
  if (vgpu_rt_launch(rt, rt->k_slab_insert, 1, 32, 0, rt->q_stream, params) == CUDA_SUCCESS &&
      spin_seq(&rt->slab_res->seq_done, seq, rt->q_stream) == 0 && rt->slab_res->slot != 0xffffffffu) {
    __sync_fetch_and_add(&rt->uva_live, 1);
    rc = 0;
  }
  ctx_leave(pushed);
  pthread_mutex_unlock(&rt->q_mu);
  return rc;
}

int vgpu_rt_slab_remove(vgpu_dev_rt *rt, CUdeviceptr dptr, uint64_t *bytes) {
  if (rt->uva_live <= 0) return 1; /* nothing recorded: not a UVA allocation */
  pthread_mutex_lock(&rt->q_mu);
  uint32_t seq = ++rt->seq;
  unsigned long long k = dptr;
  void *params[] = {&rt->slab_d, &k, &rt->slab_res_d, &seq};
  int rc = -1;
  int pushed = ctx_enter(rt);
  if (vgpu_rt_launch(rt, rt->k_slab_remove, 1, 32, 0, rt->q_stream, params) == CUDA_SUCCESS &&
      spin_seq(&rt->slab_res->seq_done, seq, rt->q_stream) == 0) {
    if (rt->slab_res->slot != 0xffffffffu) {
      *bytes = rt->slab_res->bytes;
      __sync_fetch_and_sub(&rt->uva_live, 1);
      rc = 0;
    } else {
      rc = 1;
    }
  }
  ctx_leave(pushed);
  pthread_mutex_unlock(&rt->q_mu);
  return rc;
}

static unsigned copy_grid(vgpu_dev_rt *rt, unsigned long long work_items, unsigned per_sm) {
  unsigned long long g = (unsigned long long)(rt->sm_num > 0 ? rt->sm_num : 148) * per_sm;
  if (work_items < g) g = work_items ? work_items : 1;
  return (unsigned)g;
}

CUresult vgpu_rt_clear(vgpu_dev_rt *rt, CUdeviceptr dst, size_t bytes, CUstream s) {
  unsigned long long n = bytes;
  void *params[] = {&dst, &n};
  unsigned long long tiles = (n / 16 + 2047) / 2048; /* 256 threads x 8 x 16 B per tile */
  return vgpu_rt_launch(rt, rt->k_clear, copy_grid(rt, tiles, 16), 256, 0, s, params);
}

CUresult vgpu_rt_spill(vgpu_dev_rt *rt, CUdeviceptr dst, CUdeviceptr src, size_t bytes, CUstream s) {
  unsigned long long n = bytes;
  void *params[] = {&dst, &src, &n, &rt->spill_chunk, &rt->spill_stages};
  if (((dst ^ src) & 15) != 0) /* TMA bulk copies need 16-byte congruent endpoints */
    return vgpu_rt_launch(rt, rt->k_copy_generic, copy_grid(rt, (n + 4095) / 4096, 8), 256, 0, s, params);
  unsigned long long chunks = (n + rt->spill_chunk - 1) / rt->spill_chunk;
  return vgpu_rt_launch(rt, rt->k_spill, copy_grid(rt, chunks, rt->spill_ctas_per_sm), 32,
                        rt->spill_chunk * rt->spill_stages, s, params);
}

/* ------------------------------------------------------------------ direct C-ABI (include/vgpu_b200.h) */
static vgpu_dev_rt *attached(void) {
  vgpu_boot();
  CUdevice dev;
  if (!R.cuCtxGetDevice || R.cuCtxGetDevice(&dev) != CUDA_SUCCESS) return NULL;
  vgpu_map_devices();
  return vgpu_rt_get(vgpu_host_index_of_cuda(dev), dev);
}

VGPU_EXPORT int vgpu_b200_attach(void) { return attached() ? 0 : -1; }

VGPU_EXPORT int vgpu_b200_clear(unsigned long long dst, size_t bytes, void *stream) {
  vgpu_dev_rt *rt = attached();
  if (!rt) return -1;
  return vgpu_rt_clear(rt, dst, bytes, (CUstream)stream);
}

VGPU_EXPORT int vgpu_b200_spill_copy(unsigned long long dst, unsigned long long src, size_t bytes,
                                     void *stream) {
  vgpu_dev_rt *rt = attached();
  if (!rt) return -1;
  return vgpu_rt_spill(rt, dst, src, bytes, (CUstream)stream);
}

VGPU_EXPORT int vgpu_b200_quota_eval(const void *req_, void *res_) {
  const vgpu_quota_req_t *req = (const vgpu_quota_req_t *)req_;
  vgpu_quota_res_t *res = (vgpu_quota_res_t *)res_;
  vgpu_dev_rt *rt = attached();
  if (!rt || !req || !res) return -1;
  pthread_mutex_lock(&rt->q_mu);
  memcpy(rt->q_req, req, sizeof *req);
  rt->gfx_valid = 0;      /* the hooks' cached graphics list was just overwritten */
  rt->q_req_self_set = 1; /* the caller's request is evaluated verbatim */
  int rc = vgpu_rt_quota(rt, res);
  pthread_mutex_unlock(&rt->q_mu);
  return rc;
}

VGPU_EXPORT int vgpu_b200_slab_insert(unsigned long long dptr, unsigned long long bytes) {
  vgpu_dev_rt *rt = attached();
  return rt ? vgpu_rt_slab_insert(rt, dptr, bytes) : -1;
}

VGPU_EXPORT int vgpu_b200_slab_remove(unsigned long long dptr, unsigned long long *bytes) {
  vgpu_dev_rt *rt = attached();
  uint64_t b = 0;
  int rc = rt ? vgpu_rt_slab_remove(rt, dptr, &b) : -1;
  if (bytes) *bytes = b;
  return rc;
}

static int read_lim(vgpu_dev_rt *rt, vgpu_b200_limiter_state_t *out) {
  vgpu_lim_dev_t *d = (vgpu_lim_dev_t *)malloc(sizeof *d);
  if (!d) return -1;
  int rc = R.cuMemcpyDtoH_v2(d, rt->lim_d, sizeof *d) == CUDA_SUCCESS ? 0 : -1;
  if (rc == 0 && out) {
    out->granted = d->granted;
    out->consumed = rt->lim_h->consumed;
    out->bucket = d->bucket_last;
    out->share = d->share;
    out->up_limit = d->up_limit;
    out->sys_free = d->sys_free;
    out->avg_sys_free = d->avg_sys_free;
    out->ctr_i = d->ctr_i;
    out->pre_sys_process_num = d->pre_sys_process_num;
    out->valid = d->valid;
    out->user_current = d->last_user_current;
    out->sys_current = d->last_sys_current;
    out->sm_active_pct = d->last_sm_active_pct;
    out->queue_busy_pct = d->last_queue_busy_pct;
    out->steps = d->steps;
  }
  free(d);
  return rc;
}

VGPU_EXPORT int vgpu_b200_limiter_reset(int sm_num, int max_thread_per_sm, int hard_core,
                                        int soft_core, int core_limit, int hard_limit) {
  vgpu_dev_rt *rt = attached();
  if (!rt) return -1;
  vgpu_lim_dev_t *init = (vgpu_lim_dev_t *)malloc(sizeof *init);
  if (!init) return -1;
  write_limiter_config(rt, init);
  if (sm_num > 0) {
    init->sm_num = sm_num;
    init->max_thread_per_sm = max_thread_per_sm;
    init->total_cores = (int64_t)max_thread_per_sm * (int64_t)sm_num * 32;
  }
  init->hard_core = hard_core;
  init->soft_core = soft_core;
  init->core_limit = core_limit;
  init->hard_limit = hard_limit;
  init->up_limit = hard_core;
  /* a resident governor would keep stepping the state being replaced: ask it to leave first */
  __sync_fetch_and_add(&rt->lim_h->quit, 1u);
  R.cuStreamSynchronize(rt->s_stream);
  R.cuStreamSynchronize(rt->p_stream);
  __sync_fetch_and_sub(&rt->lim_h->quit, 1u);
  rt->lim_h->gov_left_busy = 0;
  int rc = R.cuMemcpyHtoD_v2(rt->lim_d, init, sizeof *init) == CUDA_SUCCESS ? 0 : -1;
  free(init);
  rt->lim_h->consumed = 0;
  rt->lim_h->granted_mirror = 0;
  return rc;
}

VGPU_EXPORT int vgpu_b200_limiter_step(int user_current, int sys_current, int valid,
                                       int sys_process_num, vgpu_b200_limiter_state_t *out) {
  vgpu_dev_rt *rt = attached();
  if (!rt) return -1;
  vgpu_ctrl_in_t in = {user_current, sys_current, valid, sys_process_num};
  void *params[] = {&rt->lim_d, &rt->lim_h_d, &in};
  if (vgpu_rt_launch(rt, rt->k_controller, 1, 32, 0, rt->q_stream, params) != CUDA_SUCCESS) return -1;
  if (R.cuStreamSynchronize(rt->q_stream) != CUDA_SUCCESS) return -1;
  return read_lim(rt, out);
}

/* One default control step with a caller-supplied publication (vgpu_util_req_t): the fold of the
 * raw samples (reference cuda_hook.c:1044-1159) and the watcher body (:413-466) in one launch. */
VGPU_EXPORT int vgpu_b200_refill(const void *util_req, vgpu_b200_limiter_state_t *out) {
  vgpu_dev_rt *rt = attached();
  if (!rt || !util_req) return -1;
  pthread_mutex_lock(&rt->q_mu);
  memcpy(rt->u_req, util_req, sizeof(vgpu_util_req_t));
  rt->lim_h->ext_user_override = -1; /* a reading forced by an earlier vgpu_b200_sampler_run must not leak in */
  __sync_synchronize();
  uint32_t n = rt->u_req->status == VGPU_UTIL_SAMPLES ? rt->u_req->n_samples : 0;
  if (n > VGPU_MAX_PIDS) n = VGPU_MAX_PIDS;
  void *params[] = {&rt->lim_d, &rt->lim_h_d, &rt->u_req_d};
  int rc = -1;
  if (vgpu_rt_launch(rt, rt->k_refill, 1, n ? (n + 31u) & ~31u : 32u, 0, rt->q_stream, params) == CUDA_SUCCESS &&
      R.cuStreamSynchronize(rt->q_stream) == CUDA_SUCCESS)
    rc = read_lim(rt, out);
  pthread_mutex_unlock(&rt->q_mu);
  return rc;
}

/* One operation on the slab placement table (vgpu_vslab_req_t / vgpu_vslab_res_t, kernel_abi.h):
 * free-slot scan + insert, lookup + remove, or coldest-victim scan with the placement flip. */
VGPU_EXPORT int vgpu_b200_vslab_op(const void *req, void *res) {
  vgpu_dev_rt *rt = attached();
  if (!rt || !req || !res) return -1;
  pthread_mutex_lock(&rt->q_mu);
  vgpu_vslab_req_t rq = *(const vgpu_vslab_req_t *)req;
  uint32_t seq = ++rt->seq;
  if (!seq) seq = ++rt->seq;
  void *params[] = {&rt->vslab_d, &rq, &rt->vs_res_d, &seq};
  int rc = -1;
  if (vgpu_rt_launch(rt, rt->k_vslab, 1, 1024, 0, rt->q_stream, params) == CUDA_SUCCESS &&
      R.cuStreamSynchronize(rt->q_stream) == CUDA_SUCCESS && rt->vs_res->seq_done == seq) {
    *(vgpu_vslab_res_t *)res = *rt->vs_res;
    rc = 0;
  }
  pthread_mutex_unlock(&rt->q_mu);
  return rc;
}

VGPU_EXPORT int vgpu_b200_limiter_consume(long long tokens) {
  vgpu_dev_rt *rt = attached();
  if (!rt) return -1;
  __sync_fetch_and_add(&rt->lim_h->consumed, tokens);
  return 0;
}

VGPU_EXPORT int vgpu_b200_limiter_state(vgpu_b200_limiter_state_t *out) {
  vgpu_dev_rt *rt = attached();
  return rt ? read_lim(rt, out) : -1;
}

VGPU_EXPORT int vgpu_b200_sampler_run(unsigned window_us, unsigned interval_us, unsigned period_ticks,
                                      int user_override, vgpu_b200_limiter_state_t *out) {
  vgpu_dev_rt *rt = attached();
  if (!rt) return -1;
  static uint32_t epoch = 1u << 20;
  rt->lim_h->ext_user_override = user_override;
  uint32_t ep = ++epoch;
  uint32_t skipped = 0;
  void *params[] = {&rt->lim_d, &rt->lim_h_d, &window_us, &interval_us, &period_ticks, &ep, &skipped};
  unsigned grid = rt->sm_num > 0 ? (unsigned)rt->sm_num : 148u;
  if (vgpu_rt_launch(rt, rt->k_sampler, grid, 128, 0, rt->p_stream, params) != CUDA_SUCCESS) return -1;
  if (R.cuStreamSynchronize(rt->p_stream) != CUDA_SUCCESS) return -1;
  vgpu_metric_add(rt->host_index, VM_SAMPLER_LAUNCHES, 1);
  return read_lim(rt, out);
}

/* Node-agent side of the rebalance: assign a utilisation target / ceiling to the tenant(s) of
 * one GPU.  Writes VGPU_CFG_DIR/rebalance.config (the agent runs where that directory is
 * writable - the host side of the tenants' read-only config mount); every tenant's tick thread
 * applies it at its next control step.  seq must change with every new assignment.  0 / -1. */
VGPU_EXPORT int vgpu_b200_set_limits(int host_index, int up_limit, int soft_core) {
  if (host_index < 0 || host_index >= VGPU_MAX_DEVICES) return -1;
  vgpu_boot();
  int fd = open(VP(VGPU_REBALANCE_FILE), O_RDWR | O_CREAT | O_CLOEXEC, 0644);
  if (fd < 0) return -1;
  vgpu_rebalance_rec_t rec = {0, 0, 0, 0};
  off_t off = (off_t)host_index * (off_t)sizeof rec;
  ssize_t got = pread(fd, &rec, sizeof rec, off);
  (void)got;
  if (rec.magic != VGPU_REBALANCE_MAGIC) rec.seq = 0;
  rec.magic = VGPU_REBALANCE_MAGIC;
  rec.seq++;
  if (!rec.seq) rec.seq = 1;
  rec.up_limit = up_limit;
  rec.soft_core = soft_core;
  int rc = pwrite(fd, &rec, sizeof rec, off) == (ssize_t)sizeof rec ? 0 : -1;
  if (rc == 0 && lseek(fd, 0, SEEK_END) < (off_t)sizeof(vgpu_rebalance_t) && ftruncate(fd, sizeof(vgpu_rebalance_t)) != 0) rc = -1;
  close(fd);
  return rc;
}

VGPU_EXPORT int vgpu_b200_set_spill_geometry(unsigned chunk, unsigned stages, unsigned ctas_per_sm) {
  vgpu_dev_rt *rt = attached();
  if (!rt || chunk < 1024 || (chunk & 15) || stages < 2 || stages > 16 || ctas_per_sm < 1 ||
      (size_t)chunk * stages > 200 * 1024)
    return -1;
  if (R.cuFuncSetAttribute && R.cuFuncSetAttribute(rt->k_spill, 8, (int)(chunk * stages)) != CUDA_SUCCESS) return -1;
  rt->spill_chunk = chunk;
  rt->spill_stages = stages;
  rt->spill_ctas_per_sm = ctas_per_sm;
  return 0;
}

VGPU_EXPORT unsigned long long vgpu_b200_self_bytes(void) {
  vgpu_dev_rt *rt = attached();
  return rt ? rt->self_bytes : 0;
}

VGPU_EXPORT unsigned long long vgpu_b200_metric(int host_index, int which) {
  return vgpu_metric_get(host_index, which);
}

VGPU_EXPORT const char *vgpu_b200_version(void) { return "vgpu-manager_b200 0.1 (sm_100a)"; }
