/*
 * boot.c - library bring-up: logging, resolution of the real driver entry points, the
 * `dlsym` / `cuGetProcAddress` interposers and the vgpu.config contract.
 *
 * Replaces (behaviour, not code): reference library/src/loader.c:1059-1267 (real symbol
 * loading), :1290-1484 (config files), :1517-1578 (dlsym), :2054-2176 (bring-up) and
 * library/src/cuda_hook.c:1162-1314 (cuDriverGetVersion, cuInit, cuGetProcAddress[_v2]).
 *
 * Design difference: the reference resolves and re-exports ~900 driver symbols; this library
 * exports only what it intercepts, so every other symbol binds straight to the driver and the
 * reference's "ABI-conflict family" problem (cuda-helper.h:1310-1342) cannot occur.
 */
#include "vgpu_internal.h"

#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <libgen.h>
#include <link.h>
#include <regex.h>
#include <signal.h>
#include <stdarg.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>

vgpu_real_t R;
volatile unsigned vgpu_fork_epoch;
static void on_fork_child(void) { vgpu_fork_epoch++; }
vgpu_dlsym_fn vgpu_real_dlsym;
vgpu_cfg_t *G_cfg;
vgpu_smutil_t *G_smutil;
vgpu_vmem_t *G_vmem;

static vgpu_cfg_t g_env_cfg;

/* Enforcement-affecting tunables (utilisation source, periods, watchdog ...) are honoured only
 * when the limits themselves came from the tenant's environment, i.e. no control-plane config
 * is mounted; under a mounted vgpu.config the tenant's environment cannot loosen its cap. */
/* Under a mounted config the control plane can still set them: an optional VGPU_CFG_DIR/b200.tunables next to
 * vgpu.config (the same read-only mount), NAME=value per line, '#' comments.  Read once. */
static char g_tunables[4096]; /* [0] stays 0; lines, each 0-terminated, follow */
static int g_tunables_end;    /* index one past the last byte read; 0 = no file */
static void tunables_load(void) {
  int fd = open(VP(VGPU_TUNABLES_FILE), O_RDONLY | O_CLOEXEC);
  if (fd < 0) return;
  ssize_t n = read(fd, g_tunables + 1, sizeof g_tunables - 2);
  close(fd);
  if (n <= 0) return;
  g_tunables[1 + n] = 0;
  for (ssize_t i = 1; i <= n; i++)
    if (g_tunables[i] == '\n' || g_tunables[i] == '\r') g_tunables[i] = 0;
  g_tunables_end = 1 + (int)n;
}
static const char *tunable_from_file(const char *name) {
  static pthread_once_t once = PTHREAD_ONCE_INIT;
  pthread_once(&once, tunables_load);
  const size_t len = strlen(name);
  for (int i = 1; i < g_tunables_end;) {
    const char *line = g_tunables + i;
    size_t ll = strlen(line);
    if (ll > len && line[len] == '=' && !strncmp(line, name, len)) return line[len + 1] ? line + len + 1 : NULL;
    i += (int)ll + 1;
  }
  return NULL;
}

const char *vgpu_tunable(const char *name) {
  if (G_cfg && G_cfg != &g_env_cfg) return tunable_from_file(name);
  const char *s = getenv(name);
  return (s && *s) ? s : NULL;
}
static char g_driver_version[256] = "1";

const char *vgpu_path(const char *abs, char *buf, size_t cap) {
  static const char *prefix;
  static int init;
  if (!init) {
    prefix = getenv("VGPU_B200_SANDBOX");
    if (prefix && !*prefix) prefix = NULL;
    /* The tenant controls its own environment: where the control plane has mounted a config the
     * sandbox prefix must not be able to hide it (the limits would become tenant-chosen through
     * the env fallback, and locks / ledger private). */
    if (prefix && access(VGPU_CFG_FILE, F_OK) == 0) prefix = NULL;
    init = 1;
  }
  if (!prefix) return abs;
  snprintf(buf, cap, "%s%s", prefix, abs);
  return buf;
}

/* ------------------------------------------------------------------ logging */
int vgpu_log_level(void) {
  static int lvl = -1;
  if (lvl == -1) {
    int v = VL_WARNING;
    const char *s = getenv("LOGGER_LEVEL");
    if (s && *s) v = (int)strtoul(s, NULL, 10);
    if (v < VL_FATAL) v = VL_WARNING;
    if (v > VL_DETAIL) v = VL_DETAIL;
    lvl = v;
  }
  return lvl;
}

void vgpu_log_emit(int level, const char *file, int line, const char *fmt, ...) {
  static const char *names[] = {"FATAL", "ERROR", "WARNING", "INFO", "VERBOSE", "DETAIL"};
  char msg[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof msg, fmt, ap);
  va_end(ap);
  const char *base = strrchr(file, '/');
  fprintf(stderr, "[vGPU %s(%d|%" PRIuPTR "|%s:%d)]: %s\n", names[level], getpid(),
          (uintptr_t)pthread_self(), base ? base + 1 : file, line, msg);
}

const char *vgpu_cu_err(CUresult r) {
  const char *s = NULL;
  if (R.cuGetErrorString && R.cuGetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown CUDA error";
}
const char *vgpu_nv_err(nvmlReturn_t r) {
  return R.nvmlErrorString ? R.nvmlErrorString(r) : "unknown NVML error";
}

/* ------------------------------------------------------------------ real symbols */
static pthread_once_t g_dlsym_once = PTHREAD_ONCE_INIT;
static void find_real_dlsym(void) {
  /* our own `dlsym` shadows libc's, so fetch the versioned libc symbol (loader.c:1059-1090) */
  static const char *vers[] = {"GLIBC_2.2.5", "GLIBC_2.17", "GLIBC_2.34", "GLIBC_2.3",
                               "GLIBC_2.4",   "GLIBC_2.10", "GLIBC_2.18", "GLIBC_2.22", NULL};
  for (int i = 0; vers[i] && !vgpu_real_dlsym; i++)
    vgpu_real_dlsym = (vgpu_dlsym_fn)dlvsym(RTLD_NEXT, "dlsym", vers[i]);
  if (!vgpu_real_dlsym) VLOG(VL_FATAL, "unable to find the real dlsym");
}

static void read_driver_version(void) {
  /* /proc/driver/nvidia/version, first "NVRM" line, regex ([0-9]+)(\.[0-9]+)+ (loader.c:1249) */
  FILE *fp = fopen("/proc/driver/nvidia/version", "r");
  if (!fp) return;
  char *line = NULL;
  size_t cap = 0;
  while (getline(&line, &cap, fp) != -1) {
    if (strncmp(line, "NVRM", 4) != 0) continue;
    regex_t re;
    regmatch_t m[1];
    if (regcomp(&re, "([0-9]+)(\\.[0-9]+)+", REG_EXTENDED) == 0) {
      if (regexec(&re, line, 1, m, 0) == 0) {
        size_t n = (size_t)(m[0].rm_eo - m[0].rm_so);
        if (n < sizeof g_driver_version) {
          memcpy(g_driver_version, line + m[0].rm_so, n);
          g_driver_version[n] = 0;
        }
      }
      regfree(&re);
    }
    break;
  }
  free(line);
  fclose(fp);
}

static pthread_once_t g_real_once = PTHREAD_ONCE_INIT;
static void resolve_real(void) {
  pthread_once(&g_dlsym_once, find_real_dlsym);
  pthread_atfork(NULL, NULL, on_fork_child);
  read_driver_version();
  char name[512];
  snprintf(name, sizeof name, "libcuda.so.%s", g_driver_version);
  void *hc = dlopen(name, RTLD_NOW | RTLD_NODELETE);
  if (!hc) VLOG(VL_FATAL, "can't find library %s", name);
  snprintf(name, sizeof name, "libnvidia-ml.so.%s", g_driver_version);
  void *hn = dlopen(name, RTLD_NOW | RTLD_NODELETE);
  if (!hn) VLOG(VL_FATAL, "can't find library %s", name);
#define X(sym, ret, args) R.sym = (ret(*) args)vgpu_real_dlsym(hc, #sym);
  VGPU_REAL_CUDA(X)
#undef X
#define X(sym, ret, args) R.sym = (ret(*) args)vgpu_real_dlsym(hn, #sym);
  VGPU_REAL_NVML(X)
#undef X
  VLOG(VL_INFO, "resolved driver entry points (driver %s)", g_driver_version);
}

/* ------------------------------------------------------------------ vgpu.config */
static int map_ro(const char *path, size_t want, void **out) {
  if (access(path, F_OK) != 0) return 1;
  int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) {
    VLOG(VL_ERROR, "can't open %s, error %s", path, strerror(errno));
    return 1;
  }
  struct stat sb;
  int rc = 1;
  if (fstat(fd, &sb) == 0) {
    if ((size_t)sb.st_size != want) {
      VLOG(VL_ERROR, "file size mismatch: expected %zu, got %lld", want, (long long)sb.st_size);
    } else {
      void *p = mmap(NULL, want, PROT_READ, MAP_PRIVATE, fd, 0);
      if (p != MAP_FAILED) {
        *out = p;
        rc = 0;
      } else {
        VLOG(VL_ERROR, "mmap %s failed: %s", path, strerror(errno));
      }
    }
  }
  close(fd);
  return rc;
}

static int map_vmem_node(void) {
  /* first process creates + zeroes the ledger, everyone maps it shared (loader.c:1361-1412) */
  if (access(VP(VGPU_VMEM_DIR), F_OK) != 0) mkdir(VP(VGPU_VMEM_DIR), 0755);
  int created = access(VP(VGPU_VMEM_FILE), F_OK) != 0;
  int fd = open(VP(VGPU_VMEM_FILE), created ? (O_RDWR | O_CREAT | O_CLOEXEC) : (O_RDWR | O_CLOEXEC), 0644);
  if (fd < 0) {
    VLOG(VL_ERROR, "can't open %s, error %s", VP(VGPU_VMEM_FILE), strerror(errno));
    return 1;
  }
  int rc = 1;
  struct stat sb;
  if (created && ftruncate(fd, sizeof(vgpu_vmem_t)) != 0) {
    VLOG(VL_ERROR, "ftruncate failed: %s", strerror(errno));
  } else if (fstat(fd, &sb) == 0 && (created || (size_t)sb.st_size == sizeof(vgpu_vmem_t))) {
    void *p = mmap(NULL, sizeof(vgpu_vmem_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p != MAP_FAILED) {
      if (created) memset(p, 0, sizeof(vgpu_vmem_t));
      G_vmem = (vgpu_vmem_t *)p;
      rc = 0;
    }
  } else {
    VLOG(VL_ERROR, "file size mismatch: expected %zu", sizeof(vgpu_vmem_t));
  }
  close(fd);
  return rc;
}

static uint64_t iec_bytes(const char *s) { /* util.c:27-53 */
  char *end = NULL;
  double v = strtod(s, &end);
  switch (*end) {
  case 'k': case 'K': v *= 1024.0; break;
  case 'm': case 'M': v *= 1024.0 * 1024.0; break;
  case 'g': case 'G': v *= 1024.0 * 1024.0 * 1024.0; break;
  case 't': case 'T': v *= 1024.0 * 1024.0 * 1024.0 * 1024.0; break;
  default: break;
  }
  return (uint64_t)v;
}

static const char *env_for(const char *base, int idx) {
  char name[32] = {0};
  snprintf(name, sizeof name, "%s_%d", base, idx);
  const char *s = getenv(name);
  return s ? s : getenv(base);
}

static int env_true(const char *s) {
  return !strcmp(s, "true") || !strcmp(s, "TRUE") || !strcmp(s, "1");
}

static void put_field(char *dst, size_t cap, const char *name) {
  const char *s = getenv(name);
  if (!s) return;
  strncpy(dst, s, cap - 1);
  dst[cap - 1] = 0;
}

/* Build the config from the environment when the control plane did not drop a vgpu.config.
 * Field-for-field what reference loader.c:1927-2052 produces (the bytes are diffed against
 * the reference in tests/test_oracle_parity.py and tests/test_differential_stub.py). */
static void config_from_env(vgpu_cfg_t *c) {
  memset(c, 0, sizeof *c);
  const char *s = getenv("MANAGER_COMPATIBILITY_MODE");
  if (s && *s) c->compatibility_mode = (int)iec_bytes(s);
  else VLOG(VL_WARNING, "not defined env compatibility mode");
  put_field(c->pod_name, sizeof c->pod_name, "VGPU_POD_NAME");
  put_field(c->pod_namespace, sizeof c->pod_namespace, "VGPU_POD_NAMESPACE");
  put_field(c->pod_uid, sizeof c->pod_uid, "VGPU_POD_UID");
  put_field(c->container_name, sizeof c->container_name, "VGPU_CONTAINER_NAME");
  put_field(c->reg_uuid, sizeof c->reg_uuid, "MANAGER_CLIENT_REGISTER_UUID");

  char list[VGPU_UUID_LEN * VGPU_MAX_DEVICES];
  int have = 0;
  s = getenv("MANAGER_VISIBLE_DEVICES");
  if (s && *s && strlen(s) < sizeof list) {
    strcpy(list, s);
    have = 1;
  }
  if (!have) {
    /* per-index form: 16 NUL-terminated chunks; the later tokenisation only ever sees chunk 0 */
    int ok = 0;
    for (int i = 0; i < VGPU_MAX_DEVICES; i++) {
      char *slot = list + i * VGPU_UUID_LEN, name[32];
      memset(slot, 0, VGPU_UUID_LEN);
      snprintf(name, sizeof name, "MANAGER_VISIBLE_DEVICE_%d", i);
      const char *v = getenv(name);
      if (v && *v && strlen(v) < VGPU_UUID_LEN) {
        strcpy(slot, v);
        ok++;
      } else {
        strcpy(slot, VGPU_FAKE_UUID);
      }
    }
    if (!ok) {
      memset(list, 0, sizeof list);
      s = getenv("NVIDIA_VISIBLE_DEVICES");
      if (s && *s && strlen(s) < sizeof list) strcpy(list, s);
    }
  }
  char *tok[VGPU_MAX_DEVICES], *save = NULL;
  int n = 0;
  for (char *t = strtok_r(list, ",", &save); t && n < VGPU_MAX_DEVICES; t = strtok_r(NULL, ",", &save))
    tok[n++] = t;

  if ((s = getenv("VMEMORY_NODE_ENABLED"))) c->vmem_node = env_true(s);
  if ((s = getenv("EXTERNAL_SM_WATCHER_ENABLED"))) c->sm_watcher = env_true(s);

  for (int i = 0; i < n; i++) {
    if (!strcmp(tok[i], VGPU_FAKE_UUID)) continue;
    vgpu_cfg_dev_t *d = &c->devices[i];
    if (snprintf(d->uuid, VGPU_UUID_LEN, "%s", tok[i]) >= VGPU_UUID_LEN) {
      VLOG(VL_WARNING, "gpu uuid at index %d truncated", i);
      continue;
    }
    d->activate = 1;
    s = env_for("CUDA_MEM_LIMIT", i);
    if (s && *s) {
      d->total_memory = iec_bytes(s);
      d->memory_limit = 1;
    }
    int oversold = 0;
    if ((s = env_for("CUDA_MEM_OVERSOLD", i))) oversold = env_true(s);
    double ratio = 1;
    s = env_for("CUDA_MEM_RATIO", i);
    if (s && *s) ratio = atof(s);
    uint64_t real = d->total_memory;
    if (ratio > 1) {
      real /= ratio;
      d->memory_oversold = 1;
    } else {
      d->memory_oversold = oversold;
    }
    d->real_memory = real;
    int hard = 0, soft = 0;
    s = env_for("CUDA_CORE_LIMIT", i);
    if (s && *s) hard = (int)iec_bytes(s);
    if (hard > 0) {
      d->core_limit = d->hard_limit = 1;
      d->hard_core = hard;
      s = env_for("CUDA_CORE_SOFT_LIMIT", i);
      if (s && *s) soft = (int)iec_bytes(s);
      if (soft > 0 && soft > hard) {
        d->hard_limit = 0;
        d->soft_core = soft;
      }
    }
  }
}

static int publish_config(const vgpu_cfg_t *c) { /* loader.c:1453-1484 */
  if (access(VP(VGPU_ROOT_DIR), F_OK) != 0) mkdir(VP(VGPU_ROOT_DIR), 0755);
  if (access(VP(VGPU_CFG_DIR), F_OK) != 0) mkdir(VP(VGPU_CFG_DIR), 0755);
  int fd = open(VP(VGPU_CFG_FILE), O_CREAT | O_TRUNC | O_WRONLY, 0644);
  if (fd < 0) {
    VLOG(VL_ERROR, "can't open %s, error %s", VP(VGPU_CFG_FILE), strerror(errno));
    return 1;
  }
  ssize_t w = write(fd, c, sizeof *c);
  close(fd);
  if (w != (ssize_t)sizeof *c) {
    VLOG(VL_ERROR, "can't write data to %s, error %s", VP(VGPU_CFG_FILE), strerror(errno));
    return 1;
  }
  return 0;
}

/* ---- UVA-ledger hygiene (loader.c:1580-1673) ---- */
static int pid_alive(int pid) {
  if (pid <= 0) return 0;
  if (kill(pid, 0) != 0 && errno == ESRCH) return 0;
  char path[64], comm[1024], st = 0;
  int p;
  snprintf(path, sizeof path, "/proc/%d/stat", pid);
  FILE *f = fopen(path, "r");
  if (!f) return 0; /* reference: is_zombie_proc == -1 also kicks the record out */
  int ok = fscanf(f, "%d %1023s %c", &p, comm, &st) == 3;
  fclose(f);
  return ok && st != 'Z' && st != 'z';
}

static void ledger_drop(vgpu_vmem_dev_t *d, uint32_t i) {
  uint32_t n = d->processes_size;
  d->processes[i] = d->processes[n - 1];
  d->processes[n - 1].pid = 0;
  d->processes[n - 1].used = 0;
  d->processes_size = n - 1;
}

static void ledger_purge_dead(void) {
  int me = getpid();
  for (int g = 0; g < VGPU_MAX_DEVICES; g++) {
    vgpu_vmem_dev_t *d = &G_vmem->devices[g];
    if (!d->processes_size) continue;
    int fd = vgpu_vmem_lock(g, 1);
    if (fd < 0) continue;
    for (int i = (int)d->processes_size - 1; i >= 0; i--) {
      int pid = d->processes[i].pid;
      if (pid == me || !pid_alive(pid)) ledger_drop(d, (uint32_t)i);
    }
    __sync_synchronize();
    vgpu_vmem_unlock(fd, g);
  }
}

static void ledger_forget_me(void) {
  static int done;
  if (__sync_lock_test_and_set(&done, 1) || !G_vmem) return;
  int me = getpid();
  for (int g = 0; g < VGPU_MAX_DEVICES; g++) {
    vgpu_vmem_dev_t *d = &G_vmem->devices[g];
    if (!d->processes_size) continue;
    int fd = vgpu_vmem_lock(g, 1);
    if (fd < 0) continue;
    for (uint32_t i = 0; i < d->processes_size; i++)
      if (d->processes[i].pid == me) { ledger_drop(d, i); break; }
    __sync_synchronize();
    vgpu_vmem_unlock(fd, g);
  }
}

static void on_fatal_signal(int sig) {
  ledger_forget_me();
  signal(sig, SIG_DFL);
  raise(sig);
}

static void register_with_manager(void) { /* register.c:14-38, client mode only */
  pid_t child = fork();
  if (child == 0) {
    execl(VP(VGPU_CLIENT_BIN), "device-client", "--address", VGPU_ROOT_DIR "/registry/socket.sock",
          "--pod-uid", G_cfg->pod_uid, "--container-name", G_cfg->container_name,
          "--register-uuid", G_cfg->reg_uuid, (char *)NULL);
    _exit(127);
  }
  int status = 0;
  if (child < 0 || waitpid(child, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status))
    VLOG(VL_FATAL, "rpc client exit with %d", child < 0 ? -1 : WEXITSTATUS(status));
}

static pthread_mutex_t g_cfg_mu = PTHREAD_MUTEX_INITIALIZER;
static volatile unsigned g_cfg_epoch;

static void load_config(void) { /* loader.c:2054-2117; re-runs in a forked child */
  unsigned me = vgpu_fork_epoch + 1;
  if (likely(g_cfg_epoch == me)) return;
  pthread_mutex_lock(&g_cfg_mu);
  if (g_cfg_epoch != me) {
    if (!G_cfg) {
      void *p = NULL;
      if (map_ro(VP(VGPU_CFG_FILE), sizeof(vgpu_cfg_t), &p) == 0) {
        G_cfg = (vgpu_cfg_t *)p;
      } else {
        config_from_env(&g_env_cfg);
        G_cfg = &g_env_cfg;
        if (publish_config(G_cfg)) VLOG(VL_ERROR, "failed to write vgpu config file %s", VP(VGPU_CFG_FILE));
      }
    }
    if (G_cfg->sm_watcher && !G_smutil) {
      void *p = NULL;
      if (map_ro(VP(VGPU_SMUTIL_FILE), sizeof(vgpu_smutil_t), &p)) {
        pthread_mutex_unlock(&g_cfg_mu);
        VLOG(VL_FATAL, "mmap sm watcher file failed");
      }
      G_smutil = (vgpu_smutil_t *)p;
    }
    if (G_cfg->vmem_node && !G_vmem) {
      if (map_vmem_node()) {
        pthread_mutex_unlock(&g_cfg_mu);
        VLOG(VL_FATAL, "mmap vmem nodes file failed");
      }
      ledger_purge_dead();
      atexit(ledger_forget_me);
      signal(SIGTERM, on_fatal_signal);
      signal(SIGINT, on_fatal_signal);
      signal(SIGHUP, on_fatal_signal);
      signal(SIGABRT, on_fatal_signal);
    }
    if ((G_cfg->compatibility_mode & VGPU_MODE_CLIENT) == VGPU_MODE_CLIENT) register_with_manager();
    g_cfg_epoch = me;
  }
  pthread_mutex_unlock(&g_cfg_mu);
}

void vgpu_boot(void) {
  pthread_once(&g_real_once, resolve_real);
  load_config();
}

/* ------------------------------------------------------------------ dlsym interposer
 * Contract (loader.c:1517-1578): names starting with "cu"/"nvml" that we intercept resolve to
 * our entry points no matter which handle the caller passes (cudart dlopen()s libcuda and
 * dlsym()s everything); all else goes to the real dlsym.  The reference's RTLD_NEXT
 * "(tid, pointer) seen before => NULL" de-duplication is a bug we do not reproduce. */
/* RTLD_NEXT means "the first definition in the objects loaded after the CALLER".  An
 * interposed dlsym that simply forwards RTLD_NEXT answers relative to *itself* instead, which can
 * hand a later-loaded library its own symbol back (an endless loop for wrappers of the
 * "real_foo = dlsym(RTLD_NEXT, "foo")" kind; the reference papers over that by returning NULL
 * the second time a thread gets the same pointer, loader.c:1496-1539).  Walk the link map from
 * the caller's object instead. */
static void *next_after(void *caller, const char *symbol) {
  struct link_map *lm = NULL;
  Dl_info di;
  if (!caller || !dladdr1(caller, &di, (void **)&lm, RTLD_DL_LINKMAP) || !lm) return vgpu_real_dlsym(RTLD_NEXT, symbol);
  for (lm = lm->l_next; lm; lm = lm->l_next) {
    if (!lm->l_name || !lm->l_name[0]) continue; /* the main program is never "next" */
    void *h = dlopen(lm->l_name, RTLD_NOLOAD | RTLD_LAZY);
    if (!h) continue;
    void *p = vgpu_real_dlsym(h, symbol);
    dlclose(h);
    if (!p) continue;
    /* dlsym(handle) also searches the object's dependencies; accept only its own definition */
    struct link_map *owner = NULL;
    Dl_info dp;
    if (dladdr1(p, &dp, (void **)&owner, RTLD_DL_LINKMAP) && owner == lm) return p;
  }
  return NULL;
}

VGPU_EXPORT void *dlsym(void *handle, const char *symbol) {
  static __thread int depth;
  pthread_once(&g_dlsym_once, find_real_dlsym);
  if (depth > 0 || !symbol) return vgpu_real_dlsym(handle, symbol);
  if (handle == RTLD_NEXT) {
    depth++;
    void *r = next_after(__builtin_return_address(0), symbol);
    depth--;
    return r;
  }
  depth++;
  void *res = NULL;
  if (handle != RTLD_NEXT) {
    if (symbol[0] == 'c' && symbol[1] == 'u') {
      res = vgpu_lookup_cuda_hook(symbol, 0);
      if (res) vgpu_boot();
    } else if (!strncmp(symbol, "nvml", 4)) {
      res = vgpu_lookup_nvml_hook(symbol);
    }
  }
  if (!res) res = vgpu_real_dlsym(handle, symbol);
  depth--;
  return res;
}

/* ------------------------------------------------------------------ entry hooks */
VGPU_EXPORT CUresult cuDriverGetVersion(int *v) { /* cuda_hook.c:1162 - no watcher start */
  vgpu_boot();
  if (!R.cuDriverGetVersion) return CUDA_ERROR_NOT_FOUND;
  CUresult r = R.cuDriverGetVersion(v);
  if (r == CUDA_SUCCESS) vgpu_map_devices();
  return r;
}

VGPU_EXPORT CUresult cuInit(unsigned int flags) { /* cuda_hook.c:1176 */
  vgpu_boot();
  if (!R.cuInit) return CUDA_ERROR_NOT_FOUND;
  CUresult r = R.cuInit(flags);
  if (r == CUDA_SUCCESS) {
    vgpu_map_devices();
    vgpu_limiter_start();
  }
  return r;
}

static void substitute(const char *symbol, void **pfn, cuuint64_t flags, void *self) {
  if (!strcmp(symbol, "cuGetProcAddress")) {
    *pfn = self; /* keep later lookups inside the hook layer (cuda_hook.c:1213-1218) */
    return;
  }
  void *h = vgpu_lookup_cuda_hook(symbol, (flags & VCU_GET_PROC_PTDS) != 0);
  if (h) *pfn = h;
}

VGPU_EXPORT CUresult cuGetProcAddress(const char *symbol, void **pfn, int cudaVersion,
                                      cuuint64_t flags) {
  vgpu_boot();
  if (!R.cuGetProcAddress) return CUDA_ERROR_NOT_FOUND;
  CUresult r = R.cuGetProcAddress(symbol, pfn, cudaVersion, flags);
  if (r == CUDA_SUCCESS && symbol && pfn) {
    vgpu_map_devices();
    vgpu_limiter_start();
    substitute(symbol, pfn, flags, (void *)cuGetProcAddress);
  }
  return r;
}

VGPU_EXPORT CUresult cuGetProcAddress_v2(const char *symbol, void **pfn, int cudaVersion,
                                         cuuint64_t flags, void *status) {
  vgpu_boot();
  if (!R.cuGetProcAddress_v2) return CUDA_ERROR_NOT_FOUND;
  CUresult r = R.cuGetProcAddress_v2(symbol, pfn, cudaVersion, flags, status);
  if (r == CUDA_SUCCESS && symbol && pfn && *pfn) {
    vgpu_map_devices();
    vgpu_limiter_start();
    substitute(symbol, pfn, flags, (void *)cuGetProcAddress_v2);
  }
  return r;
}

VGPU_EXPORT nvmlReturn_t nvmlInitWithFlags(unsigned int flags) {
  vgpu_boot();
  return R.nvmlInitWithFlags ? R.nvmlInitWithFlags(flags) : NVML_ERROR_FUNCTION_NOT_FOUND;
}
VGPU_EXPORT nvmlReturn_t nvmlInit_v2(void) {
  vgpu_boot();
  return R.nvmlInit_v2 ? R.nvmlInit_v2() : NVML_ERROR_FUNCTION_NOT_FOUND;
}
VGPU_EXPORT nvmlReturn_t nvmlInit(void) {
  vgpu_boot();
  return R.nvmlInit ? R.nvmlInit() : NVML_ERROR_FUNCTION_NOT_FOUND;
}
