/*
 * vgpu_oracle.c - CPU restatement of the reference algorithms (see vgpu_oracle.h).
 * TEST INFRASTRUCTURE: never linked into the product library.
 * Parity: pinned differentially against oracle/_ref (the reference's own code run here).
 */
#include "vgpu_oracle.h"

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ limiter */

void orc_gpu_init(orc_gpu_t *g, int sm_num, int max_thread_per_sm) {
  g->sm_num = sm_num;
  g->max_thread_per_sm = max_thread_per_sm;
  /* cuda_hook.c:534  total = max_thread_per_sm * sm_num * FACTOR(32) */
  g->total_cores = (int64_t)max_thread_per_sm * (int64_t)sm_num * 32;
}

/* cuda_hook.c:332-367.  Constants: INCREMENT_SCALE_FACTOR 2560, MAX_UTIL_DIFF_THRESHOLD 0.5,
 * MIN_INCREMENT 5 (:33-35), error_recovery_step 10 (:289). */
int64_t orc_delta(const orc_gpu_t *g, int up_limit, int user_current, int64_t share) {
  int64_t sm = g->sm_num, thr = g->max_thread_per_sm;
  int diff = abs(up_limit - user_current);
  if (diff < 5) diff = 5;
  int64_t inc = sm * sm * thr * (int64_t)diff / 2560;
  /* the reference compares in float against the double literal 0.5 (:347) */
  if ((float)diff / (float)up_limit > 0.5) inc = inc * diff * 2 / (up_limit + 1);
  if (inc < 0 || inc > INT_MAX) inc = 10;
  if (user_current <= up_limit) {
    share = (share + inc) > g->total_cores ? g->total_cores : (share + inc);
  } else {
    share = (share - inc) < 0 ? 0 : (share - inc);
  }
  return share;
}

/* cuda_hook.c:292-306 */
int64_t orc_change_token(const orc_gpu_t *g, int64_t bucket, int64_t delta) {
  int64_t after = bucket + delta;
  if (after > g->total_cores) after = g->total_cores;
  else if (after < 0) after = 0;
  return after;
}

/* cuda_hook.c:308-330; grids is the *unsigned 32-bit* product of the grid dims converted to
 * int at the call (:1822), then widened (:317). */
int orc_rate_limiter_try(int64_t *bucket, uint32_t gx, uint32_t gy, uint32_t gz) {
  if (*bucket < 0) return 0;
  uint32_t prod = gx * gy * gz;
  int grids = (int)prod;
  *bucket -= (int64_t)grids;
  return 1;
}

void orc_watcher_init(orc_watcher_t *w, const vgpu_cfg_dev_t *cfg) {
  memset(w, 0, sizeof(*w));
  w->pre_sys_process_num = 1;
  w->up_limit = cfg->hard_core;
}

/* cuda_hook.c:413-466 for one device.  change_limit_interval 30, usage_threshold 5 (:286-290) */
void orc_watcher_step(const orc_gpu_t *g, const vgpu_cfg_dev_t *cfg, orc_watcher_t *w,
                      const orc_util_t *u, int64_t *bucket) {
  if (!cfg->core_limit) return;
  if (!u->valid) return;
  w->sys_free = 100 - u->sys_current;
  if (cfg->hard_limit) {
    if (u->sys_process_num == 1 && u->user_current < w->up_limit / 10) {
      /* jitter guard (:424-427): bucket written directly, share untouched */
      *bucket = orc_delta(g, cfg->hard_core, u->user_current, w->share);
      return;
    }
    w->share = orc_delta(g, cfg->hard_core, u->user_current, w->share);
  } else {
    if (w->pre_sys_process_num != u->sys_process_num) {
      if (w->pre_sys_process_num < u->sys_process_num) {
        w->share = (int64_t)g->max_thread_per_sm;
        w->up_limit = cfg->hard_core;
        w->i = 0;
        w->avg_sys_free = 0;
      }
      w->pre_sys_process_num = u->sys_process_num;
    }
    if (u->sys_process_num == 1) {
      w->up_limit = cfg->soft_core;
      w->share = orc_delta(g, w->up_limit, u->user_current, w->share);
    } else {
      w->i++;
      w->avg_sys_free += w->sys_free;
      if (w->i % 30 == 0) {
        if (w->avg_sys_free * 2 / 30 > 5) {
          int cand = w->up_limit + cfg->hard_core / 10;
          w->up_limit = cand > cfg->soft_core ? cfg->soft_core : cand;
        }
        w->i = 0;
      }
      w->avg_sys_free = (w->i % (30 / 2) == 0) ? 0 : w->avg_sys_free;
      w->share = orc_delta(g, w->up_limit, u->user_current, w->share);
    }
  }
  *bucket = orc_change_token(g, *bucket, w->share);
}

static int orc_valid_pct(uint32_t x) { /* GET_VALID_VALUE on an unsigned (hook.h:140) */
  return (x <= 100) ? (int)x : 0;
}

/* which membership test a compatibility mode uses; mirrors the if/else ladder shared by
 * cuda_hook.c:740-803 and :1067-1155 */
enum { ORC_SEL_NONE, ORC_SEL_PRIMARY_OPEN, ORC_SEL_OPEN_ONLY, ORC_SEL_HOST, ORC_SEL_BAD };
static int orc_mode_select(int mode, int *open_mode) {
  *open_mode = (mode & VGPU_MODE_OPEN_KERNEL) == VGPU_MODE_OPEN_KERNEL;
  if ((mode & VGPU_MODE_CLIENT) == VGPU_MODE_CLIENT) return ORC_SEL_PRIMARY_OPEN;
  if ((mode & VGPU_MODE_CGROUPV2) == VGPU_MODE_CGROUPV2) return ORC_SEL_PRIMARY_OPEN;
  if ((mode & VGPU_MODE_CGROUPV1) == VGPU_MODE_CGROUPV1) return ORC_SEL_PRIMARY_OPEN;
  if (*open_mode) return ORC_SEL_OPEN_ONLY;
  if (mode == VGPU_MODE_HOST) return ORC_SEL_HOST;
  return ORC_SEL_BAD;
}

void orc_fold_utilization(int mode, const vgpu_util_sample_t *s, uint32_t n, uint64_t checktime,
                          const uint8_t *primary, const uint8_t *local, int have_container_pids,
                          orc_util_t *u) {
  u->user_current = 0;
  u->sys_current = 0;
  if (n == 0) return;
  int open_mode, sel = orc_mode_select(mode, &open_mode);
  if (sel == ORC_SEL_BAD) return;
  /* client mode skips the whole loop when pids.config is empty (:1073) */
  if ((mode & VGPU_MODE_CLIENT) == VGPU_MODE_CLIENT && !have_container_pids) return;
  int match_primary = 0, match_open = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (s[i].ts_us < checktime) continue;
    u->valid = 1;
    int sm = orc_valid_pct(s[i].sm);
    int codec = orc_valid_pct(s[i].enc) + orc_valid_pct(s[i].dec);
    codec = codec * 85 / 100; /* CODEC_NORMALIZE hook.h:141 */
    u->sys_current += sm + codec;
    switch (sel) {
    case ORC_SEL_HOST:
      u->user_current += sm + codec;
      break;
    case ORC_SEL_OPEN_ONLY:
      if (local[i]) u->user_current += sm + codec;
      break;
    default:
      if (!match_open && primary[i]) {
        match_primary = 1;
        u->user_current += sm + codec;
      } else if (!match_primary && open_mode && local[i]) {
        match_open = 1;
        u->user_current += sm + codec;
      }
    }
  }
}

int orc_balance_batches(int device_count, int sm_watcher, int *start, int *end) {
  if (device_count <= 0) return 0;
  int batch_size = sm_watcher ? VGPU_MAX_DEVICES / 2 : 4;
  int batch_count = (device_count + batch_size - 1) / batch_size;
  int base = device_count / batch_count, rem = device_count % batch_count, cur = 0;
  for (int i = 0; i < batch_count; i++) {
    int sz = base + (i < rem ? 1 : 0);
    start[i] = cur;
    end[i] = cur + sz;
    cur += sz;
  }
  return batch_count;
}

/* ------------------------------------------------------------------ memory */

uint64_t orc_accumulate_used(int mode, const vgpu_proc_t *p, uint32_t n, const uint8_t *primary,
                             const uint8_t *local) {
  uint64_t used = 0;
  if (n == 0) return 0;
  int open_mode, sel = orc_mode_select(mode, &open_mode);
  int match_primary = 0, match_open = 0;
  for (uint32_t i = 0; i < n; i++) {
    switch (sel) {
    case ORC_SEL_HOST:
      used += p[i].used_bytes;
      break;
    case ORC_SEL_OPEN_ONLY:
      if (local[i]) used += p[i].used_bytes;
      break;
    case ORC_SEL_PRIMARY_OPEN:
      if (!match_open && primary[i]) {
        match_primary = 1;
        used += p[i].used_bytes;
      } else if (!match_primary && open_mode && local[i]) {
        match_open = 1;
        used += p[i].used_bytes;
      }
      break;
    default:
      break;
    }
  }
  return used;
}

uint64_t orc_used_memory(int mode, const vgpu_proc_t *comp, uint32_t nc, const uint8_t *cprim,
                         const uint8_t *cloc, const vgpu_proc_t *gfx, uint32_t ng,
                         const uint8_t *gprim, const uint8_t *gloc) {
  uint64_t used = orc_accumulate_used(mode, comp, nc, cprim, cloc);
  /* graphics entries whose pid already appears in the compute list are dropped, order of
   * the survivors preserved (cuda_hook.c:868-887) */
  vgpu_proc_t *uniq = (vgpu_proc_t *)malloc(sizeof(vgpu_proc_t) * (ng ? ng : 1));
  uint8_t *up = (uint8_t *)malloc(ng ? ng : 1), *ul = (uint8_t *)malloc(ng ? ng : 1);
  uint32_t k = 0;
  if (!uniq || !up || !ul) abort(); /* test infrastructure: no partial answers */
  for (uint32_t i = 0; i < ng; i++) {
    int seen = 0;
    for (uint32_t j = 0; j < nc; j++)
      if (gfx[i].pid == comp[j].pid) { seen = 1; break; }
    if (seen) continue;
    uniq[k] = gfx[i];
    up[k] = gprim ? gprim[i] : 0;
    ul[k] = gloc ? gloc[i] : 0;
    k++;
  }
  used += orc_accumulate_used(mode, uniq, k, up, ul);
  free(uniq); free(up); free(ul);
  return used;
}

int orc_memory_path(const vgpu_cfg_dev_t *cfg, uint64_t used, uint64_t vmem, uint64_t request,
                    int allow_uva) {
  if (!cfg->memory_limit) return ORC_PATH_GPU; /* load_limited_memory_view returned 0 */
  if ((used + vmem + request) > cfg->total_memory) return ORC_PATH_OOM; /* wraps like size_t */
  if (allow_uva && cfg->memory_oversold && (used + request) > cfg->real_memory) return ORC_PATH_UVA;
  return ORC_PATH_GPU;
}

void orc_nvml_meminfo(const vgpu_cfg_dev_t *cfg, uint64_t used, uint64_t vmem, uint64_t *total,
                      uint64_t *out_used, uint64_t *out_free) {
  uint64_t t = cfg->total_memory, tu = used + vmem;
  *total = t;
  *out_used = tu >= t ? t : tu;
  *out_free = *total - *out_used;
}

void orc_cu_meminfo(const vgpu_cfg_dev_t *cfg, uint64_t used, uint64_t vmem, int real_ok,
                    uint64_t real_total, uint64_t *free_out, uint64_t *total_out) {
  uint64_t configured = cfg->total_memory, actual;
  if (cfg->memory_oversold) actual = configured;
  else actual = (real_ok && real_total > 0 && real_total < configured) ? real_total : configured;
  *total_out = actual;
  *free_out = (used + vmem) >= actual ? 0 : (actual - used - vmem);
}

static uint64_t orc_array_base(int format) { /* cuda_hook.c:1546-1569 (bits) */
  switch (format) {
  case 0x01: case 0x08: return 8;             /* (UN)SIGNED_INT8  */
  case 0x02: case 0x09: case 0x10: return 16; /* (UN)SIGNED_INT16, HALF */
  case 0x03: case 0x0a: case 0x20: return 32; /* (UN)SIGNED_INT32, FLOAT */
  default: return 32;
  }
}
uint64_t orc_array_request(int format, uint64_t channels, uint64_t h, uint64_t w) {
  return orc_array_base(format) * channels * h * w;
}
uint64_t orc_array3d_request(int format, uint64_t channels, uint64_t h, uint64_t w, uint64_t d) {
  return orc_array_base(format) * channels * h * w * d;
}
uint64_t orc_pitch_guess(uint64_t width_bytes, uint32_t elem) {
  return (((width_bytes - 1) / elem) + 1) * elem;
}

int orc_ledger_add(vgpu_vmem_dev_t *d, int pid, uint64_t bytes) {
  uint32_t n = d->processes_size;
  for (uint32_t i = 0; i < n; i++)
    if (d->processes[i].pid == pid) { d->processes[i].used += bytes; return 0; }
  if (n >= VGPU_MAX_PIDS) return -1;
  d->processes[n].pid = pid;
  d->processes[n].used = bytes;
  d->processes_size++;
  return 0;
}

void orc_ledger_sub(vgpu_vmem_dev_t *d, int pid, uint64_t bytes) {
  for (uint32_t i = 0; i < d->processes_size; i++)
    if (d->processes[i].pid == pid) {
      d->processes[i].used = d->processes[i].used >= bytes ? d->processes[i].used - bytes : 0;
      break;
    }
}

uint64_t orc_ledger_sum(const vgpu_vmem_dev_t *d) {
  uint64_t s = 0;
  for (uint32_t i = 0; i < d->processes_size; i++) s += d->processes[i].used;
  return s;
}

void orc_ledger_rm_pid(vgpu_vmem_dev_t *d, int pid) {
  uint32_t n = d->processes_size;
  for (uint32_t i = 0; i < n; i++)
    if (d->processes[i].pid == pid) {
      d->processes[i] = d->processes[n - 1];
      d->processes[n - 1].pid = 0;
      d->processes[n - 1].used = 0;
      d->processes_size--;
      return;
    }
}

void orc_ledger_purge(vgpu_vmem_dev_t *d, int self_pid, const uint8_t *alive) {
  uint32_t n = d->processes_size;
  for (int i = (int)n - 1; i >= 0; i--) {
    int kick = d->processes[i].pid == self_pid || !alive[i];
    if (kick) {
      d->processes[i] = d->processes[n - 1];
      d->processes[n - 1].pid = 0;
      d->processes[n - 1].used = 0;
      d->processes_size--;
      n--;
    }
  }
}

/* ------------------------------------------------------------------ env -> config */

uint64_t orc_iec_to_bytes(const char *s) {
  char *end = NULL;
  double v = strtod(s, &end);
  switch (*end) {
  case 'K': case 'k': v *= 1024UL; break;
  case 'M': case 'm': v *= 1024UL * 1024UL; break;
  case 'G': case 'g': v *= 1024UL * 1024UL * 1024UL; break;
  case 'T': case 't': v *= 1024UL * 1024UL * 1024UL * 1024UL; break;
  default: break;
  }
  return (uint64_t)v;
}

static const char *orc_env_idx(orc_getenv_fn ge, void *ctx, const char *base, int idx) {
  char name[32] = {0};
  snprintf(name, sizeof(name), "%s_%d", base, idx);
  const char *s = ge(name, ctx);
  if (!s) s = ge(base, ctx);
  return s;
}

static int orc_truthy(const char *s) {
  return strcmp(s, "true") == 0 || strcmp(s, "TRUE") == 0 || strcmp(s, "1") == 0;
}

static void orc_copy_field(char *dst, size_t cap, const char *src) {
  if (!src) return;
  strncpy(dst, src, cap - 1);
  dst[cap - 1] = '\0';
}

void orc_config_from_env(orc_getenv_fn ge, void *ctx, vgpu_cfg_t *cfg) {
  memset(cfg, 0, sizeof(*cfg));
  const char *s = ge("MANAGER_COMPATIBILITY_MODE", ctx);
  if (s && s[0]) cfg->compatibility_mode = (int)orc_iec_to_bytes(s);
  orc_copy_field(cfg->pod_name, sizeof(cfg->pod_name), ge("VGPU_POD_NAME", ctx));
  orc_copy_field(cfg->pod_namespace, sizeof(cfg->pod_namespace), ge("VGPU_POD_NAMESPACE", ctx));
  orc_copy_field(cfg->pod_uid, sizeof(cfg->pod_uid), ge("VGPU_POD_UID", ctx));
  orc_copy_field(cfg->container_name, sizeof(cfg->container_name), ge("VGPU_CONTAINER_NAME", ctx));
  orc_copy_field(cfg->reg_uuid, sizeof(cfg->reg_uuid), ge("MANAGER_CLIENT_REGISTER_UUID", ctx));

  /* uuid list: MANAGER_VISIBLE_DEVICES, else per-index MANAGER_VISIBLE_DEVICE_<i> (of which
   * only chunk 0 survives the later strtok, loader.c:1959-1987), else NVIDIA_VISIBLE_DEVICES */
  char uuids[VGPU_UUID_LEN * VGPU_MAX_DEVICES];
  int have = 0;
  s = ge("MANAGER_VISIBLE_DEVICES", ctx);
  if (s && s[0] && strlen(s) < sizeof(uuids)) {
    snprintf(uuids, sizeof(uuids), "%s", s);
    have = 1;
  }
  if (!have) {
    int ok = 0;
    for (int i = 0; i < VGPU_MAX_DEVICES; i++) {
      char *slot = &uuids[i * VGPU_UUID_LEN];
      memset(slot, 0, VGPU_UUID_LEN);
      char name[32] = {0};
      snprintf(name, sizeof(name), "MANAGER_VISIBLE_DEVICE_%d", i);
      const char *v = ge(name, ctx);
      if (v && v[0] && strlen(v) < VGPU_UUID_LEN) {
        snprintf(slot, VGPU_UUID_LEN, "%s", v);
        ok++;
      } else {
        /* (a too-long value is first written truncated, then ...) */
        strncpy(slot, VGPU_FAKE_UUID, VGPU_UUID_LEN - 1);         /* ... overwritten by the fake */
        slot[VGPU_UUID_LEN - 1] = '\0';
      }
    }
    if (!ok) {
      memset(uuids, 0, sizeof(uuids));
      s = ge("NVIDIA_VISIBLE_DEVICES", ctx);
      if (s && s[0] && strlen(s) < sizeof(uuids)) snprintf(uuids, sizeof(uuids), "%s", s);
    }
  }

  char *tok[VGPU_MAX_DEVICES];
  int n = 0;
  char *save = NULL;
  for (char *t = strtok_r(uuids, ",", &save); t && n < VGPU_MAX_DEVICES;
       t = strtok_r(NULL, ",", &save))
    tok[n++] = t;

  s = ge("VMEMORY_NODE_ENABLED", ctx);
  if (s) cfg->vmem_node = orc_truthy(s);
  s = ge("EXTERNAL_SM_WATCHER_ENABLED", ctx);
  if (s) cfg->sm_watcher = orc_truthy(s);

  for (int i = 0; i < n; i++) {
    if (strcmp(tok[i], VGPU_FAKE_UUID) == 0) continue;
    vgpu_cfg_dev_t *d = &cfg->devices[i];
    if (snprintf(d->uuid, VGPU_UUID_LEN, "%s", tok[i]) >= VGPU_UUID_LEN) continue;
    d->activate = 1;
    s = orc_env_idx(ge, ctx, "CUDA_MEM_LIMIT", i);
    if (s && s[0]) {
      d->total_memory = orc_iec_to_bytes(s);
      d->memory_limit = 1;
    } else {
      d->memory_limit = 0;
    }
    int oversold = 0;
    s = orc_env_idx(ge, ctx, "CUDA_MEM_OVERSOLD", i);
    if (s) oversold = orc_truthy(s);
    double ratio = 1;
    s = orc_env_idx(ge, ctx, "CUDA_MEM_RATIO", i);
    if (s && s[0]) ratio = atof(s);
    uint64_t real = d->total_memory;
    if (ratio > 1) {
      real /= ratio; /* size_t /= double, as in loader.c:2019 */
      d->memory_oversold = 1;
    } else {
      d->memory_oversold = oversold;
    }
    d->real_memory = real;
    int hard = 0, soft = 0;
    s = orc_env_idx(ge, ctx, "CUDA_CORE_LIMIT", i);
    if (s && s[0]) hard = (int)orc_iec_to_bytes(s);
    if (hard > 0) {
      d->core_limit = 1;
      d->hard_limit = 1;
      d->hard_core = hard;
      s = orc_env_idx(ge, ctx, "CUDA_CORE_SOFT_LIMIT", i);
      if (s && s[0]) soft = (int)orc_iec_to_bytes(s);
      if (soft > 0 && soft > hard) {
        d->hard_limit = 0;
        d->soft_core = soft;
      }
    } else {
      d->core_limit = 0;
      d->hard_limit = 0;
    }
  }
}
