/*
 * config.c - device index maps, cross-process locks and container-pid predicates.
 *
 * Behavioural contract:
 *   index maps        reference library/src/loader.c:1675-1822, :2119-2164
 *   per-GPU file lock reference library/src/lock.c:39-107 (F_SETLK polled every 10 ms, 5 s cap)
 *   byte-range locks  reference library/src/lock.c:108-232 (F_SETLKW at offsetof(lock_byte))
 *   pid predicates    reference library/src/cuda_hook.c:617-733, library/src/util.c:221-331
 *
 * Design difference: the reference takes one global pthread mutex on *every* launch to map
 * cuda device -> host index (loader.c:1769); here the maps are lock-free arrays filled once.
 */
#include "vgpu_internal.h"

#include <errno.h>
#include <fcntl.h>
#include <limits.h>
#include <signal.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <time.h>

static volatile int g_cuda2host[VGPU_MAX_DEVICES];
static volatile int g_cuda2nvml[VGPU_MAX_DEVICES];
static volatile int g_nvml2host[VGPU_MAX_DEVICES];
static nvmlDevice_t g_host2nvml[VGPU_MAX_DEVICES];
static volatile unsigned g_map_epoch; /* == vgpu_fork_epoch + 1 once initialised */
static pthread_mutex_t g_map_mu = PTHREAD_MUTEX_INITIALIZER;

/* the reference compares getpid() on every call (loader.c:1808-1822); a pthread_atfork child
 * handler bumping an epoch gives the same fork awareness without a syscall per launch */
static inline void maps_reset_if_forked(void) {
  unsigned want = vgpu_fork_epoch + 1;
  if (likely(g_map_epoch == want)) return;
  pthread_mutex_lock(&g_map_mu);
  if (g_map_epoch != want) {
    for (int i = 0; i < VGPU_MAX_DEVICES; i++) {
      g_cuda2host[i] = -1;
      g_cuda2nvml[i] = -1;
      if (g_map_epoch == 0) g_nvml2host[i] = -1;
    }
    g_map_epoch = want;
  }
  pthread_mutex_unlock(&g_map_mu);
}

static int host_index_of_uuid(const char *uuid) {
  for (int i = 0; i < VGPU_MAX_DEVICES; i++)
    if (G_cfg->devices[i].activate && strcmp(G_cfg->devices[i].uuid, uuid) == 0) return i;
  return -1;
}

int vgpu_host_index_of_cuda(CUdevice dev) {
  if (unlikely(dev < 0 || dev >= VGPU_MAX_DEVICES)) {
    VLOG(VL_ERROR, "invalid cuda index %d", dev);
    return -1;
  }
  maps_reset_if_forked();
  int h = g_cuda2host[dev];
  if (likely(h >= 0)) return h;
  CUuuid u;
  CUresult r = R.cuDeviceGetUuid_v2 ? R.cuDeviceGetUuid_v2(&u, dev)
               : R.cuDeviceGetUuid  ? R.cuDeviceGetUuid(&u, dev)
                                    : CUDA_ERROR_NOT_FOUND;
  if (r != CUDA_SUCCESS) {
    VLOG(VL_VERBOSE, "cuDeviceGetUuid can't get uuid on cuda device %d, return %d", dev, r);
    return -1;
  }
  const uint8_t *b = (const uint8_t *)u.bytes;
  char s[VGPU_UUID_LEN];
  snprintf(s, sizeof s, "GPU-%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x",
           b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8], b[9], b[10], b[11], b[12], b[13],
           b[14], b[15]);
  h = host_index_of_uuid(s);
  if (h >= 0) {
    g_cuda2host[dev] = h;
    VLOG(VL_VERBOSE, "cuda device %d => host device %d", dev, h);
  }
  return h;
}

int vgpu_host_index_of_nvml(nvmlDevice_t dev) {
  unsigned int ni = 0;
  if (!R.nvmlDeviceGetIndex || R.nvmlDeviceGetIndex(dev, &ni) != NVML_SUCCESS) return -1;
  if (ni >= VGPU_MAX_DEVICES) {
    VLOG(VL_ERROR, "invalid nvml index %u", ni);
    return -1;
  }
  maps_reset_if_forked();
  int h = g_nvml2host[ni];
  if (likely(h >= 0)) return h;
  char uuid[VGPU_UUID_LEN];
  if (!R.nvmlDeviceGetUUID || R.nvmlDeviceGetUUID(dev, uuid, VGPU_UUID_LEN) != NVML_SUCCESS) return -1;
  h = host_index_of_uuid(uuid);
  if (h >= 0) {
    g_host2nvml[h] = dev;
    g_nvml2host[ni] = h;
    VLOG(VL_VERBOSE, "nvml device %u => host device %d", ni, h);
  }
  return h;
}

int vgpu_nvml_index_of_cuda(CUdevice dev) {
  if (dev < 0 || dev >= VGPU_MAX_DEVICES) return -1;
  maps_reset_if_forked();
  int n = g_cuda2nvml[dev];
  if (n >= 0) return n;
  int h = vgpu_host_index_of_cuda(dev);
  if (h < 0) return -1;
  for (int i = 0; i < VGPU_MAX_DEVICES; i++)
    if (g_nvml2host[i] == h) {
      g_cuda2nvml[dev] = i;
      return i;
    }
  return -1;
}

nvmlDevice_t vgpu_nvml_handle_of_host(int h) {
  return (h >= 0 && h < VGPU_MAX_DEVICES) ? g_host2nvml[h] : NULL;
}

static pthread_once_t g_nvml_map_once = PTHREAD_ONCE_INIT;
static void build_nvml_map(void) { /* loader.c:2119-2164 */
  nvmlReturn_t rt = R.nvmlInitWithFlags ? R.nvmlInitWithFlags(0)
                    : R.nvmlInit_v2     ? R.nvmlInit_v2()
                    : R.nvmlInit        ? R.nvmlInit()
                                        : NVML_ERROR_FUNCTION_NOT_FOUND;
  if (rt) VLOG(VL_FATAL, "nvmlInit failed, return: %d, str: %s", rt, vgpu_nv_err(rt));
  unsigned int n = 0;
  rt = R.nvmlDeviceGetCount      ? R.nvmlDeviceGetCount(&n)
       : R.nvmlDeviceGetCount_v2 ? R.nvmlDeviceGetCount_v2(&n)
                                 : NVML_ERROR_FUNCTION_NOT_FOUND;
  if (rt) VLOG(VL_FATAL, "nvmlDeviceGetCount call failed, return: %d, str: %s", rt, vgpu_nv_err(rt));
  maps_reset_if_forked();
  for (unsigned int i = 0; i < n; i++) {
    nvmlDevice_t d;
    rt = R.nvmlDeviceGetHandleByIndex_v2 ? R.nvmlDeviceGetHandleByIndex_v2(i, &d)
         : R.nvmlDeviceGetHandleByIndex  ? R.nvmlDeviceGetHandleByIndex(i, &d)
                                         : NVML_ERROR_FUNCTION_NOT_FOUND;
    if (rt) {
      VLOG(VL_ERROR, "nvmlDeviceGetHandleByIndex call failed, nvml device: %u, return: %d", i, rt);
      continue;
    }
    vgpu_host_index_of_nvml(d);
  }
}

void vgpu_map_devices(void) { pthread_once(&g_nvml_map_once, build_nvml_map); }

/* ------------------------------------------------------------------ locks */
/* Per-GPU whole-file write lock (lock.c:47-99).  The reference opens, locks, unlocks and closes
 * the file around every allocation (5 system calls); here the descriptor stays open per process,
 * so taking the lock is one fcntl and dropping it another.  POSIX record locks belong to the
 * process, not the descriptor, so the semantics towards other processes are identical; inside
 * the process the runtime's q_mu serialises.  The value returned is a token (fd + 1 would do;
 * the fd itself is kept for the old call sites) that vgpu_unlock_gpu understands. */
static int g_lock_fd[VGPU_MAX_DEVICES];       /* fd + 1 */
static unsigned g_lock_epoch[VGPU_MAX_DEVICES];
static pthread_mutex_t g_lock_open_mu = PTHREAD_MUTEX_INITIALIZER;

static int lock_file_fd(int h) {
  if (g_lock_fd[h] && g_lock_epoch[h] == vgpu_fork_epoch + 1) return g_lock_fd[h] - 1;
  pthread_mutex_lock(&g_lock_open_mu);
  if (!g_lock_fd[h] || g_lock_epoch[h] != vgpu_fork_epoch + 1) {
    /* (a forked child keeps using the inherited descriptor number, but re-opens so that the two
     * processes do not share one open file description) */
    if (access(VP(VGPU_LOCK_DIR), F_OK) != 0) mkdir(VP(VGPU_LOCK_DIR), 0755);
    char raw[64];
    snprintf(raw, sizeof raw, VGPU_LOCK_FMT, h);
    int fd = open(VP(raw), O_RDWR | O_CREAT | O_CLOEXEC, 0644);
    if (fd >= 0) {
      g_lock_fd[h] = fd + 1;
      g_lock_epoch[h] = vgpu_fork_epoch + 1;
    }
  }
  int out = g_lock_fd[h] && g_lock_epoch[h] == vgpu_fork_epoch + 1 ? g_lock_fd[h] - 1 : -1;
  pthread_mutex_unlock(&g_lock_open_mu);
  return out;
}

int vgpu_lock_gpu(int h) {
  if (h < 0 || h >= VGPU_MAX_DEVICES) {
    VLOG(VL_ERROR, "invalid device index %d", h);
    return -1;
  }
  struct timespec t0, now, nap = {0, 10 * 1000 * 1000};
  int timed = 0;
  for (;;) {
    int fd = lock_file_fd(h);
    if (fd >= 0) {
      struct flock fl = {.l_type = F_WRLCK, .l_whence = SEEK_SET, .l_start = 0, .l_len = 0};
      if (fcntl(fd, F_SETLK, &fl) == 0) return fd;
    }
    if (!timed) {
      clock_gettime(CLOCK_MONOTONIC, &t0);
      timed = 1;
    }
    clock_gettime(CLOCK_MONOTONIC, &now);
    long ms = (now.tv_sec - t0.tv_sec) * 1000 + (now.tv_nsec - t0.tv_nsec) / 1000000;
    if (ms >= 5000) {
      vgpu_metric_add(h, VM_LOCK_TIMEOUT, 1);
      VLOG(VL_ERROR, "lock timeout for device %d", h);
      return -1;
    }
    nanosleep(&nap, NULL);
  }
}

void vgpu_unlock_gpu(int fd) {
  if (fd < 0) return;
  struct flock fl = {.l_type = F_UNLCK, .l_whence = SEEK_SET, .l_start = 0, .l_len = 0};
  fcntl(fd, F_SETLK, &fl); /* the descriptor stays open (see above) */
}

static int range_lock(const char *path, int oflags, off_t off, short type) {
  int fd = open(path, oflags | O_CLOEXEC, 0644);
  if (fd < 0) {
    VLOG(VL_ERROR, "failed to open shared file %s: %s", path, strerror(errno));
    return -1;
  }
  struct flock fl = {.l_type = type, .l_whence = SEEK_SET, .l_start = off, .l_len = 1};
  if (fcntl(fd, F_SETLKW, &fl) == -1) {
    VLOG(VL_ERROR, "fcntl lock failed on %s: %s", path, strerror(errno));
    close(fd);
    return -1;
  }
  return fd;
}

static void range_unlock(int fd, off_t off) {
  if (fd < 0) return;
  struct flock fl = {.l_type = F_UNLCK, .l_whence = SEEK_SET, .l_start = off, .l_len = 1};
  fcntl(fd, F_SETLK, &fl);
  close(fd);
}

int vgpu_vmem_lock(int h, int write) {
  if (h < 0 || h >= VGPU_MAX_DEVICES) return -1;
  off_t off = (off_t)offsetof(vgpu_vmem_t, devices) + (off_t)h * (off_t)sizeof(vgpu_vmem_dev_t) +
              (off_t)offsetof(vgpu_vmem_dev_t, lock_byte);
  return write ? range_lock(VP(VGPU_VMEM_FILE), O_RDWR | O_CREAT, off, F_WRLCK)
               : range_lock(VP(VGPU_VMEM_FILE), O_RDONLY, off, F_RDLCK);
}
void vgpu_vmem_unlock(int fd, int h) {
  if (h < 0 || h >= VGPU_MAX_DEVICES) return;
  range_unlock(fd, (off_t)h * (off_t)sizeof(vgpu_vmem_dev_t) + (off_t)offsetof(vgpu_vmem_dev_t, lock_byte));
}
int vgpu_smutil_rdlock(int h) {
  if (h < 0 || h >= VGPU_MAX_DEVICES) return -1;
  return range_lock(VP(VGPU_SMUTIL_FILE), O_RDONLY,
                    (off_t)h * (off_t)sizeof(vgpu_smutil_dev_t) + (off_t)offsetof(vgpu_smutil_dev_t, lock_byte),
                    F_RDLCK);
}
void vgpu_smutil_unlock(int fd, int h) {
  if (h < 0 || h >= VGPU_MAX_DEVICES) return;
  range_unlock(fd, (off_t)h * (off_t)sizeof(vgpu_smutil_dev_t) + (off_t)offsetof(vgpu_smutil_dev_t, lock_byte));
}

/* ------------------------------------------------------------------ container pid predicates */
static int cgroup_value(const char *path, const char *key, char *out, size_t cap) {
  /* lines are "<id>:<controllers>:<path>" (cuda_hook.c:617-642) */
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  char line[256];
  int rc = -1;
  while (fgets(line, sizeof line, f)) {
    char *c1 = strchr(line, ':');
    if (!c1) continue;
    char *c2 = strchr(c1 + 1, ':');
    if (!c2) continue;
    *c2 = 0;
    char *k = c1 + 1;
    while (*k == ' ' || *k == '\t') k++;
    if (strcmp(k, key) != 0) continue;
    char *v = c2 + 1;
    while (*v == ' ' || *v == '\t') v++;
    size_t n = strlen(v);
    while (n && (v[n - 1] == '\n' || v[n - 1] == '\t')) n--;
    if (n >= cap) n = cap - 1;
    memcpy(out, v, n);
    out[n] = 0;
    rc = 0;
    break;
  }
  fclose(f);
  return rc;
}

static int in_cgroup_v1(unsigned pid) {
  if (!pid) return 0;
  char hp[128], mine[256], theirs[256];
  snprintf(hp, sizeof hp, VGPU_HOSTPROC_CGROUP_FMT, (int)pid);
  if (cgroup_value("/proc/self/cgroup", "memory", mine, sizeof mine)) return 0;
  if (cgroup_value(VP(hp), "memory", theirs, sizeof theirs)) return 0;
  return strstr(theirs, mine) != NULL;
}

static int in_cgroup_v2(unsigned pid) {
  if (!pid) return 0;
  char hp[128], line[FILENAME_MAX];
  snprintf(hp, sizeof hp, VGPU_HOSTPROC_CGROUP_FMT, (int)pid);
  FILE *f = fopen(VP(hp), "r");
  if (!f) return 0;
  int hit = 0;
  while (!hit && fgets(line, sizeof line, f)) {
    size_t n = strlen(line);
    if (n && line[n - 1] == '\n') line[n - 1] = 0;
    hit = strcmp(line, "0::/") == 0;
  }
  fclose(f);
  return hit;
}

static int same_ns(unsigned pid, const char *ns) {
  char a[128], b[128];
  struct stat sa, sb;
  snprintf(a, sizeof a, "/proc/%u/ns/%s", pid, ns);
  snprintf(b, sizeof b, "/proc/self/ns/%s", ns);
  return stat(a, &sa) == 0 && stat(b, &sb) == 0 && sa.st_ino == sb.st_ino;
}

static int is_local_gpu_pid(unsigned pid) { /* cuda_hook.c:715-733 */
  if (!pid) return 0;
  if (kill((pid_t)pid, 0) != 0 && errno == ESRCH) return 0;
  if (!same_ns(pid, "mnt") || !same_ns(pid, "cgroup")) return 0;
  char path[64], line[1024];
  snprintf(path, sizeof path, "/proc/%u/maps", pid);
  FILE *f = fopen(path, "r");
  if (!f) return 0;
  int hit = 0;
  while (!hit && fgets(line, sizeof line, f)) {
    /* 6th space-separated token is the mapped path (util.c:269-303) */
    char *save = NULL, *t = strtok_r(line, " ", &save);
    for (int k = 1; t && k < 6; k++) t = strtok_r(NULL, " ", &save);
    hit = t && strstr(t, "nvidia");
  }
  fclose(f);
  return hit;
}

static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

static int read_container_pids(int *pids, int cap) { /* util.c:221-267 */
  if (access(VP(VGPU_PIDS_FILE), F_OK) != 0) return 0;
  FILE *f = fopen(VP(VGPU_PIDS_FILE), "r");
  if (!f) return 0;
  char line[32];
  int n = 0;
  while (n < cap && fgets(line, sizeof line, f)) {
    char *end;
    long v = strtol(line, &end, 10);
    if (end == line || (*end != '\n' && *end != 0)) continue;
    if (v <= 0 || v > INT_MAX) continue;
    pids[n++] = (int)v;
  }
  fclose(f);
  if (n > 0) qsort(pids, (size_t)n, sizeof(int), cmp_int);
  return n;
}

static int pid_flags(const uint32_t *pids, uint32_t n, uint8_t *flags, int fatal_if_unregistered);
void vgpu_pid_flags(const uint32_t *pids, uint32_t n, uint8_t *flags) { pid_flags(pids, n, flags, 1); }
int vgpu_pid_flags_util(const uint32_t *pids, uint32_t n, uint8_t *flags) { return pid_flags(pids, n, flags, 0); }

static int pid_flags(const uint32_t *pids, uint32_t n, uint8_t *flags, int fatal_if_unregistered) {
  int mode = G_cfg->compatibility_mode;
  int open_mode = (mode & VGPU_MODE_OPEN_KERNEL) == VGPU_MODE_OPEN_KERNEL;
  memset(flags, 0, n);
  if (n == 0) return 1;
  if ((mode & VGPU_MODE_CLIENT) == VGPU_MODE_CLIENT) {
    static __thread int cpids[VGPU_MAX_PIDS];
    int cn = read_container_pids(cpids, VGPU_MAX_PIDS);
    if (cn == 0) {
      if (fatal_if_unregistered) VLOG(VL_FATAL, "unable to find registered container process");
      return 0;
    }
    for (uint32_t i = 0; i < n; i++) {
      int key = (int)pids[i];
      if (pids[i] && bsearch(&key, cpids, (size_t)cn, sizeof(int), cmp_int)) flags[i] |= VGPU_FLAG_PRIMARY;
    }
  } else if ((mode & VGPU_MODE_CGROUPV2) == VGPU_MODE_CGROUPV2) {
    for (uint32_t i = 0; i < n; i++)
      if (in_cgroup_v2(pids[i])) flags[i] |= VGPU_FLAG_PRIMARY;
  } else if ((mode & VGPU_MODE_CGROUPV1) == VGPU_MODE_CGROUPV1) {
    for (uint32_t i = 0; i < n; i++)
      if (in_cgroup_v1(pids[i])) flags[i] |= VGPU_FLAG_PRIMARY;
  } else if (!open_mode && mode != VGPU_MODE_HOST) {
    VLOG(VL_FATAL, "unsupported environment compatibility mode: %d", mode);
  }
  /* the local test must be known for every pid: once the open-kernel branch has latched, the
   * reference evaluates it even for pids that also pass the primary test (cuda_hook.c:751-759) */
  if (open_mode)
    for (uint32_t i = 0; i < n; i++)
      if (is_local_gpu_pid(pids[i])) flags[i] |= VGPU_FLAG_LOCAL;
  return 1;
}

/* ------------------------------------------------------------------ own-footprint registry
 * Every process that brings the device runtime up owns a few MiB of HBM (module + limiter
 * state).  NVML charges them to the tenant, the reference has no such footprint, so the quota
 * kernel removes them again.  One process only knows its own share; siblings in the same
 * container publish theirs in a small table next to the per-GPU lock file (which is already a
 * cross-process shared directory), keyed by container identity.  All accesses happen while the
 * per-GPU lock is held. */
typedef struct { int32_t pid; uint32_t key; uint64_t bytes; } self_rec_t;
#define SELF_RECS 1023 /* record 0 of the file is the header {generation} */

static uint32_t container_key(void) {
  uint32_t h = 2166136261u;
  const char *parts[2] = {G_cfg->pod_uid, G_cfg->container_name};
  for (int p = 0; p < 2; p++)
    for (const char *s = parts[p]; *s; s++) h = (h ^ (uint8_t)*s) * 16777619u;
  return h ? h : 1;
}

static int g_self_fd[VGPU_MAX_DEVICES];          /* fd + 1, kept open */
static uint64_t g_self_gen[VGPU_MAX_DEVICES];    /* generation the cached total was computed at */
static uint64_t g_self_total[VGPU_MAX_DEVICES];
static time_t g_self_checked[VGPU_MAX_DEVICES];  /* last liveness sweep */
static unsigned g_self_epoch[VGPU_MAX_DEVICES];

uint64_t vgpu_self_registry(int h, uint64_t publish_bytes, int publish) {
  if (h < 0 || h >= VGPU_MAX_DEVICES) return publish_bytes;
  if (g_self_epoch[h] != vgpu_fork_epoch + 1) { /* a forked child starts with nothing cached */
    g_self_fd[h] = 0;
    g_self_gen[h] = 0;
    g_self_epoch[h] = vgpu_fork_epoch + 1;
  }
  if (!g_self_fd[h]) {
    char raw[64];
    snprintf(raw, sizeof raw, VGPU_LOCK_DIR "/vgpu_%d.b200", h);
    int fd = open(VP(raw), O_RDWR | O_CREAT | O_CLOEXEC, 0644);
    if (fd < 0) return publish_bytes;
    g_self_fd[h] = fd + 1;
  }
  int fd = g_self_fd[h] - 1;
  /* fast path: nothing was (un)registered since the cached sum, and the last liveness sweep is
   * recent - one 8-byte pread */
  uint64_t gen = 0;
  time_t now = time(NULL);
  if (publish <= 0 && pread(fd, &gen, sizeof gen, 0) == (ssize_t)sizeof gen && gen == g_self_gen[h] && gen != 0 &&
      now - g_self_checked[h] < 2)
    return g_self_total[h];

  static __thread self_rec_t tab[SELF_RECS + 1];
  memset(tab, 0, sizeof tab);
  ssize_t got = pread(fd, tab, sizeof tab, 0);
  (void)got;
  memcpy(&gen, &tab[0], sizeof gen);
  uint32_t key = container_key();
  int me = getpid(), mine = -1, free_slot = -1, dirty = 0;
  uint64_t total = 0;
  for (int i = 1; i <= SELF_RECS; i++) {
    if (tab[i].pid == 0) { if (free_slot < 0) free_slot = i; continue; }
    if (tab[i].pid == me && tab[i].key == key) { mine = i; continue; }
    if (tab[i].key != key) continue;
    if (kill(tab[i].pid, 0) != 0 && errno == ESRCH) { memset(&tab[i], 0, sizeof tab[i]); dirty = 1; if (free_slot < 0) free_slot = i; continue; }
    total += tab[i].bytes;
  }
  if (publish > 0) {
    int slot = mine >= 0 ? mine : free_slot;
    if (slot >= 0) { tab[slot].pid = me; tab[slot].key = key; tab[slot].bytes = publish_bytes; dirty = 1; mine = slot; }
  }
  total += mine >= 0 ? tab[mine].bytes : publish_bytes;
  if ((dirty || gen == 0) && publish >= 0) { /* publish < 0: a reader that does not hold the GPU lock never writes */
    gen++;
    memcpy(&tab[0], &gen, sizeof gen);
    ssize_t w = pwrite(fd, tab, sizeof tab, 0);
    (void)w;
  }
  g_self_gen[h] = gen;
  g_self_total[h] = total;
  g_self_checked[h] = now;
  return total;
}
