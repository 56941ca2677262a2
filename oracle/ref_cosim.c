/*
 * ref_cosim.c - runs the REFERENCE's own limiter / accounting code under a line protocol so
 * that oracle/vgpu_oracle.c (and through it the CUDA kernels) can be pinned against it.
 * TEST INFRASTRUCTURE.  Built by oracle/Makefile into oracle/_ref/ref_cosim from the reference
 * sources where they lie (/root/reference/library); nothing of the reference is copied.
 *
 * Technique: this translation unit #includes the reference's src/cuda_hook.c, so its `static`
 * functions (delta, change_token, rate_limiter, utilization_watcher ...) and per-device statics
 * are reachable.  nanosleep/gettimeofday are macro-redirected to a virtual clock: the watcher
 * thread parks in its per-iteration nanosleep and is released one iteration at a time, with
 * the utilisation samples for that iteration scripted through fake NVML table entries.
 *
 * Protocol (stdin -> stdout, one reply line per command):
 *   delta SM THR UP USER SHARE                -> share'
 *   token SM THR BUCKET DELTA                 -> bucket'
 *   rate BUCKET GX GY GZ                      -> bucket'           (BUCKET must be >= 0)
 *   setenv K V | unsetenv K                   -> ok
 *   envcfg                                    -> 3696 hex chars (the 1848-byte config)
 *   used MODE NC pid:bytes.. NG pid:bytes..   -> used              (get_used_gpu_memory_by_device)
 *   winit MODE HARD SOFT CORE_LIMIT HARD_LIMIT SM THR -> ok
 *   wset BUCKET                               -> ok                (host consumption between ticks)
 *   wstep NPROC NS pid:sm:enc:dec:age_ms..    -> share bucket up_limit valid user sys
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "include/hook.h"
#include "include/cuda-helper.h"
#include "include/metrics.h"
#include "include/nvml-helper.h"

/* ---- virtual clock + lock-step handshake ---- */
static pthread_mutex_t vt_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t vt_cv = PTHREAD_COND_INITIALIZER;
static pthread_t vt_watcher;
static int vt_have_watcher, vt_parked, vt_go;
static unsigned long long vt_now_us = 1700000000ull * 1000000ull;

static int vt_nanosleep(const struct timespec *req, struct timespec *rem) {
  (void)rem;
  if (vt_have_watcher && pthread_equal(pthread_self(), vt_watcher)) {
    pthread_mutex_lock(&vt_mu);
    vt_parked = 1;
    pthread_cond_broadcast(&vt_cv);
    while (!vt_go) pthread_cond_wait(&vt_cv, &vt_mu);
    vt_go = 0;
    vt_parked = 0;
    vt_now_us += (unsigned long long)req->tv_sec * 1000000ull + (unsigned long long)req->tv_nsec / 1000ull;
    pthread_mutex_unlock(&vt_mu);
  }
  return 0;
}
static int vt_gettimeofday(struct timeval *tv, void *tz) {
  (void)tz;
  tv->tv_sec = (time_t)(vt_now_us / 1000000ull);
  tv->tv_usec = (suseconds_t)(vt_now_us % 1000000ull);
  return 0;
}

#define nanosleep vt_nanosleep
#define gettimeofday vt_gettimeofday
#include "src/cuda_hook.c"
#undef nanosleep
#undef gettimeofday

/* ---- scripted fake driver entries ---- */
#define MAXS 1024
static nvmlProcessInfo_t s_comp[MAXS], s_gfx[MAXS];
static unsigned s_ncomp, s_ngfx;
static nvmlProcessUtilizationSample_t s_samples[MAXS];
static unsigned s_nsamples;
static int s_nproc;

static const unsigned char k_uuid_bytes[16] = {0x11, 0x11, 0x11, 0x11, 0x11, 0x11, 0x11, 0x11,
                                               0x11, 0x11, 0x11, 0x11, 0x11, 0x11, 0x11, 0x11};
static CUresult fake_cuDeviceGetUuid(CUuuid *u, CUdevice d) { (void)d; memcpy(u->bytes, k_uuid_bytes, 16); return CUDA_SUCCESS; }
static CUresult fake_cuGetErrorString(CUresult r, const char **s) { (void)r; *s = "fake"; return CUDA_SUCCESS; }
static const char *fake_nvmlErrorString(nvmlReturn_t r) { (void)r; return "fake"; }
static nvmlReturn_t fake_compute(nvmlDevice_t d, unsigned *n, nvmlProcessInfo_t *out) {
  (void)d;
  /* the watcher only uses the count; the memory path uses the records */
  unsigned k = vt_have_watcher && pthread_equal(pthread_self(), vt_watcher) ? (unsigned)s_nproc : s_ncomp;
  if (!(vt_have_watcher && pthread_equal(pthread_self(), vt_watcher))) memcpy(out, s_comp, s_ncomp * sizeof *out);
  *n = k;
  return NVML_SUCCESS;
}
static nvmlReturn_t fake_graphics(nvmlDevice_t d, unsigned *n, nvmlProcessInfo_t *out) {
  (void)d;
  if (vt_have_watcher && pthread_equal(pthread_self(), vt_watcher)) { *n = 0; return NVML_SUCCESS; }
  memcpy(out, s_gfx, s_ngfx * sizeof *out);
  *n = s_ngfx;
  return NVML_SUCCESS;
}
static nvmlReturn_t fake_procutil(nvmlDevice_t d, nvmlProcessUtilizationSample_t *out, unsigned *n, unsigned long long since) {
  (void)d; (void)since;
  if (s_nsamples == 0) { *n = 0; return NVML_ERROR_NOT_FOUND; }
  memcpy(out, s_samples, s_nsamples * sizeof *out);
  *n = s_nsamples;
  return NVML_SUCCESS;
}

extern resource_data_t vgpu_config_temp;
extern int init_g_vgpu_config_by_env();
static resource_data_t my_cfg;

static void install_fakes(void) {
  cuda_library_entry[CUDA_ENTRY_ENUM(cuDeviceGetUuid_v2)].fn_ptr = fake_cuDeviceGetUuid;
  cuda_library_entry[CUDA_ENTRY_ENUM(cuDeviceGetUuid)].fn_ptr = fake_cuDeviceGetUuid;
  cuda_library_entry[CUDA_ENTRY_ENUM(cuGetErrorString)].fn_ptr = fake_cuGetErrorString;
  nvml_library_entry[NVML_ENTRY_ENUM(nvmlErrorString)].fn_ptr = fake_nvmlErrorString;
  nvml_library_entry[NVML_ENTRY_ENUM(nvmlDeviceGetComputeRunningProcesses)].fn_ptr = fake_compute;
  nvml_library_entry[NVML_ENTRY_ENUM(nvmlDeviceGetGraphicsRunningProcesses)].fn_ptr = fake_graphics;
  nvml_library_entry[NVML_ENTRY_ENUM(nvmlDeviceGetProcessUtilization)].fn_ptr = fake_procutil;
}

static void base_cfg(int mode) {
  memset(&my_cfg, 0, sizeof my_cfg);
  my_cfg.compatibility_mode = mode;
  snprintf(my_cfg.devices[0].uuid, UUID_BUFFER_SIZE, "GPU-11111111-1111-1111-1111-111111111111");
  my_cfg.devices[0].activate = 1;
  g_vgpu_config = &my_cfg;
}

/* loader.c:1084 references this private glibc symbol; shared objects may leave it undefined,
 * an executable may not.  Never called (dlvsym succeeds first). */
void *_dl_sym(void *h, const char *n, void *w) { (void)h; (void)n; (void)w; return NULL; }

static void *watcher_entry(void *arg) { return utilization_watcher(arg); }

int main(void) {
  char line[1 << 16];
  install_fakes();
  setvbuf(stdout, NULL, _IOLBF, 0);
  while (fgets(line, sizeof line, stdin)) {
    char *save = NULL, *cmd = strtok_r(line, " \n", &save);
    if (!cmd) continue;
    if (!strcmp(cmd, "delta")) {
      int sm = atoi(strtok_r(NULL, " \n", &save)), thr = atoi(strtok_r(NULL, " \n", &save));
      int up = atoi(strtok_r(NULL, " \n", &save)), user = atoi(strtok_r(NULL, " \n", &save));
      long long share = atoll(strtok_r(NULL, " \n", &save));
      g_sm_num[0] = sm; g_max_thread_per_sm[0] = thr;
      g_total_cuda_cores[0] = (int64_t)thr * (int64_t)sm * FACTOR;
      printf("%lld\n", (long long)delta(up, user, share, 0));
    } else if (!strcmp(cmd, "token")) {
      int sm = atoi(strtok_r(NULL, " \n", &save)), thr = atoi(strtok_r(NULL, " \n", &save));
      long long b = atoll(strtok_r(NULL, " \n", &save)), d = atoll(strtok_r(NULL, " \n", &save));
      g_total_cuda_cores[0] = (int64_t)thr * (int64_t)sm * FACTOR;
      g_cur_cuda_cores[0] = b;
      change_token(d, 0);
      printf("%lld\n", (long long)g_cur_cuda_cores[0]);
    } else if (!strcmp(cmd, "rate")) {
      long long b = atoll(strtok_r(NULL, " \n", &save));
      unsigned gx = strtoul(strtok_r(NULL, " \n", &save), NULL, 10), gy = strtoul(strtok_r(NULL, " \n", &save), NULL, 10),
               gz = strtoul(strtok_r(NULL, " \n", &save), NULL, 10);
      base_cfg(0);
      my_cfg.devices[0].core_limit = 1;
      g_cur_cuda_cores[0] = b;
      rate_limiter(gx * gy * gz, 1, 0); /* same unsigned product -> int conversion as :1822 */
      printf("%lld\n", (long long)g_cur_cuda_cores[0]);
    } else if (!strcmp(cmd, "setenv")) {
      char *k = strtok_r(NULL, " \n", &save), *v = strtok_r(NULL, "\n", &save);
      setenv(k, v ? v : "", 1);
      printf("ok\n");
    } else if (!strcmp(cmd, "unsetenv")) {
      unsetenv(strtok_r(NULL, " \n", &save));
      printf("ok\n");
    } else if (!strcmp(cmd, "envcfg")) {
      memset(&vgpu_config_temp, 0, sizeof vgpu_config_temp);
      init_g_vgpu_config_by_env();
      const unsigned char *p = (const unsigned char *)g_vgpu_config;
      for (size_t i = 0; i < sizeof(resource_data_t); i++) printf("%02x", p[i]);
      printf("\n");
    } else if (!strcmp(cmd, "used")) {
      int mode = atoi(strtok_r(NULL, " \n", &save));
      base_cfg(mode);
      s_ncomp = (unsigned)atoi(strtok_r(NULL, " \n", &save));
      for (unsigned i = 0; i < s_ncomp; i++) {
        char *t = strtok_r(NULL, " \n", &save);
        unsigned pid; unsigned long long b;
        sscanf(t, "%u:%llu", &pid, &b);
        s_comp[i].pid = pid; s_comp[i].usedGpuMemory = b;
      }
      s_ngfx = (unsigned)atoi(strtok_r(NULL, " \n", &save));
      for (unsigned i = 0; i < s_ngfx; i++) {
        char *t = strtok_r(NULL, " \n", &save);
        unsigned pid; unsigned long long b;
        sscanf(t, "%u:%llu", &pid, &b);
        s_gfx[i].pid = pid; s_gfx[i].usedGpuMemory = b;
      }
      size_t used = 0;
      get_used_gpu_memory_by_device(&used, (nvmlDevice_t)0x1);
      printf("%zu\n", used);
    } else if (!strcmp(cmd, "winit")) {
      int mode = atoi(strtok_r(NULL, " \n", &save));
      base_cfg(mode);
      my_cfg.devices[0].hard_core = atoi(strtok_r(NULL, " \n", &save));
      my_cfg.devices[0].soft_core = atoi(strtok_r(NULL, " \n", &save));
      my_cfg.devices[0].core_limit = atoi(strtok_r(NULL, " \n", &save));
      my_cfg.devices[0].hard_limit = atoi(strtok_r(NULL, " \n", &save));
      g_sm_num[0] = atoi(strtok_r(NULL, " \n", &save));
      g_max_thread_per_sm[0] = atoi(strtok_r(NULL, " \n", &save));
      g_total_cuda_cores[0] = (int64_t)g_max_thread_per_sm[0] * (int64_t)g_sm_num[0] * FACTOR;
      g_cur_cuda_cores[0] = 0;
      if (vt_have_watcher) { printf("err one watcher per process\n"); continue; }
      batches[0].start_index = 0; batches[0].end_index = 1; batches[0].batch_code = 0;
      pthread_mutex_lock(&vt_mu);
      vt_have_watcher = 1;
      pthread_create(&vt_watcher, NULL, watcher_entry, &batches[0]);
      while (!vt_parked) pthread_cond_wait(&vt_cv, &vt_mu);
      pthread_mutex_unlock(&vt_mu);
      printf("ok\n");
    } else if (!strcmp(cmd, "wset")) {
      g_cur_cuda_cores[0] = atoll(strtok_r(NULL, " \n", &save));
      printf("ok\n");
    } else if (!strcmp(cmd, "wstep")) {
      s_nproc = atoi(strtok_r(NULL, " \n", &save));
      s_nsamples = (unsigned)atoi(strtok_r(NULL, " \n", &save));
      /* the watcher adds its sleep (80 ms) to the clock before sampling */
      unsigned long long t_sample = vt_now_us + 80000ull;
      for (unsigned i = 0; i < s_nsamples; i++) {
        char *t = strtok_r(NULL, " \n", &save);
        unsigned pid, sm, enc, dec; long long age_ms;
        sscanf(t, "%u:%u:%u:%u:%lld", &pid, &sm, &enc, &dec, &age_ms);
        memset(&s_samples[i], 0, sizeof s_samples[i]);
        s_samples[i].pid = pid; s_samples[i].smUtil = sm; s_samples[i].encUtil = enc; s_samples[i].decUtil = dec;
        s_samples[i].timeStamp = t_sample - (unsigned long long)(age_ms * 1000);
      }
      pthread_mutex_lock(&vt_mu);
      vt_go = 1;
      pthread_cond_broadcast(&vt_cv);
      while (vt_go || !vt_parked) pthread_cond_wait(&vt_cv, &vt_mu);
      pthread_mutex_unlock(&vt_mu);
      printf("%lld %lld %d %d %d %d\n", (long long)shares[0], (long long)g_cur_cuda_cores[0], up_limits[0],
             top_results[0].valid, top_results[0].user_current, top_results[0].sys_current);
    } else {
      printf("err unknown command %s\n", cmd);
    }
  }
  return 0;
}
