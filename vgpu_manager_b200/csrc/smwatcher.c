/*
 * smwatcher.c - node-level producer of /etc/vgpu-manager/watcher/sm_util.config.
 *
 * Replaces:
 *   reference pkg/device/manager/watcher.go:53-184   (SMUtilWatcherStart / smWatcherSingleDevice)
 *   reference pkg/config/watcher/sm_watcher.go:34-131 (file layout, byte-range locks)
 *   reference pkg/config/watcher/batch.go             (BalanceBatches - one loop here, see below)
 *
 * Consumers (unchanged): the interception library's external-watcher mode
 * (library/src/cuda_hook.c:1009-1042, csrc/limiter.c refresh_from_external_watcher) and the Go
 * device-monitor.  Every tenant then reads one shared file instead of issuing its own NVML
 * ioctls every 80 ms.
 *
 * Behaviour kept from the Go producer:
 *   - file created 0644 and sized to sizeof(device_util_t) = 1 311 232 bytes if absent / wrong size
 *   - per device, in NVML index order, every pass:
 *       compute + graphics process lists (24-byte nvmlProcessInfo v2/v3 records, <= 1024),
 *       process utilisation samples newer than now - 1 s (32-byte records, <= 1024),
 *       lastSeenTimeStamp = now - 1 s (microseconds);
 *     the samples (and their count) are only replaced when the NVML query succeeded
 *   - the device record is updated under an F_WRLCK byte-range lock on its lock_byte (F_SETLKW)
 *   - MIG-enabled devices are skipped
 *   - one pass over all devices takes ~80 ms (the Go code sleeps 80 ms / batch size per device)
 *
 * The Go code spreads devices over goroutines in batches; a node has at most 16 GPUs and one
 * NVML sweep costs well under a millisecond per device, so this is a single loop.
 *
 * Plain C, libc only; NVML is dlopen'ed like the library does (libnvidia-ml.so.1).
 *
 * On-device readings (--source device|mixed): tenants whose limiter runs on an on-device
 * utilisation signal publish their reading every control period in VGPU_LOCK_DIR/vgpu_<i>.readings
 * (include/vgpu_contract.h; csrc/limiter.c publish_device_reading).  With --source device the samples of a device
 * are built from those readings alone - nvmlDeviceGetProcessUtilization is never called; with --source mixed NVML
 * is asked and a tenant's own reading replaces NVML's sample for its pid.  Readings are keyed by (pid-namespace
 * inode, pid inside it); NVML's host pids are resolved through /proc/<pid>/ns/pid and the NSpid line of
 * /proc/<pid>/status.  The process lists (memory accounting) always come from NVML.
 *
 *   vgpu-smwatcher [--file PATH] [--passes N] [--period-ms 80] [--nvml PATH] [--verbose]
 *                  [--source nvml|device|mixed] [--readings-dir DIR]
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../../include/vgpu_contract.h"

typedef int nvml_ret;
typedef void *nvml_dev;
typedef struct {
  nvml_ret (*init)(void);
  nvml_ret (*shutdown)(void);
  nvml_ret (*count)(unsigned *);
  nvml_ret (*by_index)(unsigned, nvml_dev *);
  nvml_ret (*mig_mode)(nvml_dev, unsigned *, unsigned *);
  /* 24-byte records (v2/v3 ABI) when available, else the 16-byte v1 ABI */
  nvml_ret (*compute24)(nvml_dev, unsigned *, vgpu_proc_v2_t *);
  nvml_ret (*graphics24)(nvml_dev, unsigned *, vgpu_proc_v2_t *);
  nvml_ret (*compute16)(nvml_dev, unsigned *, vgpu_proc_t *);
  nvml_ret (*graphics16)(nvml_dev, unsigned *, vgpu_proc_t *);
  nvml_ret (*proc_util)(nvml_dev, vgpu_util_sample_t *, unsigned *, unsigned long long);
} nvml_api;

enum { SRC_NVML = 0, SRC_DEVICE = 1, SRC_MIXED = 2 };
static int g_source = SRC_NVML;
static const char *g_readings_dir = VGPU_LOCK_DIR;

static volatile sig_atomic_t g_stop;
static void on_signal(int s) { (void)s; g_stop = 1; }

static void *sym(void *h, const char *a, const char *b, const char *c) {
  void *p = a ? dlsym(h, a) : NULL;
  if (!p && b) p = dlsym(h, b);
  if (!p && c) p = dlsym(h, c);
  return p;
}

static int load_nvml(const char *path, nvml_api *n) {
  void *h = dlopen(path ? path : "libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "vgpu-smwatcher: %s\n", dlerror()); return -1; }
  memset(n, 0, sizeof *n);
  n->init = (nvml_ret(*)(void))sym(h, "nvmlInit_v2", "nvmlInit", NULL);
  n->shutdown = (nvml_ret(*)(void))sym(h, "nvmlShutdown", NULL, NULL);
  n->count = (nvml_ret(*)(unsigned *))sym(h, "nvmlDeviceGetCount_v2", "nvmlDeviceGetCount", NULL);
  n->by_index = (nvml_ret(*)(unsigned, nvml_dev *))sym(h, "nvmlDeviceGetHandleByIndex_v2", "nvmlDeviceGetHandleByIndex", NULL);
  n->mig_mode = (nvml_ret(*)(nvml_dev, unsigned *, unsigned *))sym(h, "nvmlDeviceGetMigMode", NULL, NULL);
  n->compute24 = (nvml_ret(*)(nvml_dev, unsigned *, vgpu_proc_v2_t *))sym(h, "nvmlDeviceGetComputeRunningProcesses_v3",
                                                                           "nvmlDeviceGetComputeRunningProcesses_v2", NULL);
  n->graphics24 = (nvml_ret(*)(nvml_dev, unsigned *, vgpu_proc_v2_t *))sym(h, "nvmlDeviceGetGraphicsRunningProcesses_v3",
                                                                            "nvmlDeviceGetGraphicsRunningProcesses_v2", NULL);
  n->compute16 = (nvml_ret(*)(nvml_dev, unsigned *, vgpu_proc_t *))sym(h, "nvmlDeviceGetComputeRunningProcesses", NULL, NULL);
  n->graphics16 = (nvml_ret(*)(nvml_dev, unsigned *, vgpu_proc_t *))sym(h, "nvmlDeviceGetGraphicsRunningProcesses", NULL, NULL);
  n->proc_util = (nvml_ret(*)(nvml_dev, vgpu_util_sample_t *, unsigned *, unsigned long long))sym(
      h, "nvmlDeviceGetProcessUtilization", NULL, NULL);
  if (!n->init || !n->count || !n->by_index || !(n->compute24 || n->compute16) || (!n->proc_util && g_source != SRC_DEVICE)) {
    fprintf(stderr, "vgpu-smwatcher: NVML library lacks a required entry point\n");
    return -1;
  }
  return 0;
}

/* PrepareDeviceUtilFile (sm_watcher.go): make sure the file exists with the exact size */
static vgpu_smutil_t *map_file(const char *path) {
  char dir[4096];
  snprintf(dir, sizeof dir, "%s", path);
  char *slash = strrchr(dir, '/');
  if (slash && slash != dir) {
    *slash = 0;
    for (char *p = dir + 1;; p++) { /* mkdir -p */
      if (*p == '/' || *p == 0) {
        char c = *p;
        *p = 0;
        if (mkdir(dir, 0755) != 0 && errno != EEXIST) { perror(dir); return NULL; }
        *p = c;
        if (!c) break;
      }
    }
  }
  int fd = open(path, O_RDWR | O_CREAT | O_CLOEXEC, 0644);
  if (fd < 0) { perror(path); return NULL; }
  struct stat st;
  if (fstat(fd, &st) != 0 || (size_t)st.st_size != sizeof(vgpu_smutil_t)) {
    if (ftruncate(fd, 0) != 0 || ftruncate(fd, (off_t)sizeof(vgpu_smutil_t)) != 0) { perror("ftruncate"); close(fd); return NULL; }
  }
  void *m = mmap(NULL, sizeof(vgpu_smutil_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  return m == MAP_FAILED ? NULL : (vgpu_smutil_t *)m;
}

static int wlock(const char *path, int ordinal) {
  int fd = open(path, O_RDWR | O_CREAT | O_CLOEXEC, 0644);
  if (fd < 0) return -1;
  struct flock fl = {.l_type = F_WRLCK, .l_whence = SEEK_SET,
                     .l_start = (off_t)(offsetof(vgpu_smutil_t, devices) + (size_t)ordinal * sizeof(vgpu_smutil_dev_t) +
                                        offsetof(vgpu_smutil_dev_t, lock_byte)),
                     .l_len = 1};
  if (fcntl(fd, F_SETLKW, &fl) == -1) { close(fd); return -1; }
  return fd;
}
static void unlock(int fd, int ordinal) {
  if (fd < 0) return;
  struct flock fl = {.l_type = F_UNLCK, .l_whence = SEEK_SET,
                     .l_start = (off_t)(offsetof(vgpu_smutil_t, devices) + (size_t)ordinal * sizeof(vgpu_smutil_dev_t) +
                                        offsetof(vgpu_smutil_dev_t, lock_byte)),
                     .l_len = 1};
  fcntl(fd, F_SETLK, &fl);
  close(fd);
}

/* one NVML process list as 24-byte records; returns the count or -1 */
static int list_procs(const nvml_api *n, nvml_dev d, int graphics, vgpu_proc_v2_t *out) {
  unsigned cnt = VGPU_MAX_PIDS;
  nvml_ret (*f24)(nvml_dev, unsigned *, vgpu_proc_v2_t *) = graphics ? n->graphics24 : n->compute24;
  if (f24) {
    nvml_ret r = f24(d, &cnt, out);
    if (r != 0) return -1;
    return (int)(cnt > VGPU_MAX_PIDS ? VGPU_MAX_PIDS : cnt);
  }
  static vgpu_proc_t tmp[VGPU_MAX_PIDS];
  nvml_ret (*f16)(nvml_dev, unsigned *, vgpu_proc_t *) = graphics ? n->graphics16 : n->compute16;
  if (!f16 || f16(d, &cnt, tmp) != 0) return -1;
  if (cnt > VGPU_MAX_PIDS) cnt = VGPU_MAX_PIDS;
  for (unsigned i = 0; i < cnt; i++) {
    out[i].pid = tmp[i].pid;
    out[i]._pad = 0;
    out[i].used_bytes = tmp[i].used_bytes;
    out[i].gi = out[i].ci = 0xFFFFFFFFu; /* "not a MIG instance", what NVML itself reports */
  }
  return (int)cnt;
}

static unsigned long long now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  return (unsigned long long)ts.tv_sec * 1000000ull + (unsigned long long)ts.tv_nsec / 1000ull;
}

/* ------------------------------------------------------------------ on-device readings */
static const vgpu_readings_t *readings_of(int i) {
  static const vgpu_readings_t *map[VGPU_MAX_DEVICES];
  if (map[i]) return map[i];
  char path[4096];
  snprintf(path, sizeof path, "%s/vgpu_%d.readings", g_readings_dir, i);
  int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) return NULL; /* no tenant has published on this GPU yet: looked for again next pass */
  struct stat st;
  void *m = (fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(vgpu_readings_t))
                ? mmap(NULL, sizeof(vgpu_readings_t), PROT_READ, MAP_SHARED, fd, 0) : MAP_FAILED;
  close(fd);
  if (m == MAP_FAILED) return NULL;
  return map[i] = (const vgpu_readings_t *)m;
}

/* the key a tenant publishes under, for one of NVML's (host) pids: its pid namespace and its pid in there */
static uint64_t owner_key_of(unsigned pid) {
  char path[64], line[256];
  struct stat ns;
  snprintf(path, sizeof path, "/proc/%u/ns/pid", pid);
  if (stat(path, &ns) != 0) return 0;
  snprintf(path, sizeof path, "/proc/%u/status", pid);
  FILE *f = fopen(path, "r");
  if (!f) return 0;
  unsigned inner = pid;
  while (fgets(line, sizeof line, f))
    if (!strncmp(line, "NSpid:", 6)) { /* outermost ... innermost */
      for (char *tok = strtok(line + 6, " \t\n"); tok; tok = strtok(NULL, " \t\n")) inner = (unsigned)strtoul(tok, NULL, 10);
      break;
    }
  fclose(f);
  return ((uint64_t)ns.st_ino << 32) | inner;
}

/* samples for the compute processes that published a reading since `since_us`; returns their number */
static unsigned samples_from_readings(int i, const vgpu_proc_v2_t *compute, int nc, unsigned long long since_us, vgpu_util_sample_t *out) {
  const vgpu_readings_t *R = readings_of(i);
  if (!R) return 0;
  unsigned n = 0;
  for (int p = 0; p < nc && n < VGPU_MAX_PIDS; p++) {
    uint64_t key = owner_key_of(compute[p].pid);
    if (!key) continue;
    for (int k = 0; k < VGPU_READINGS_SLOTS; k++) {
      if (R->slots[k].owner != key) continue;
      unsigned long long ts = R->slots[k].ts_us;
      __sync_synchronize();
      unsigned sm = R->slots[k].sm_pct;
      if (ts >= since_us && R->slots[k].owner == key) {
        memset(&out[n], 0, sizeof out[n]);
        out[n].pid = compute[p].pid;
        out[n].ts_us = ts;
        out[n].sm = sm > 100 ? 100 : sm;
        n++;
      }
      break;
    }
  }
  return n;
}

/* smWatcherSingleDevice (watcher.go:128-184) */
static int publish_device(const nvml_api *n, vgpu_smutil_t *file, const char *path, int i, nvml_dev d) {
  static vgpu_proc_v2_t compute[VGPU_MAX_PIDS], graphics[VGPU_MAX_PIDS];
  static vgpu_util_sample_t samples[VGPU_MAX_PIDS];
  if (n->mig_mode) {
    unsigned cur = 0, pend = 0;
    if (n->mig_mode(d, &cur, &pend) == 0 && cur == 1) return 0;
  }
  int nc = list_procs(n, d, 0, compute);
  if (nc < 0) return 0; /* logged-and-skipped in the Go code */
  int ng = n->graphics24 || n->graphics16 ? list_procs(n, d, 1, graphics) : 0;
  if (ng < 0) return 0;
  unsigned long long last_ts = now_us() - 1000000ull;
  unsigned ns = VGPU_MAX_PIDS;
  nvml_ret sr = 0;
  if (g_source == SRC_DEVICE) {
    /* the tenants' own readings, nothing else; none fresh = NVML's NOT_FOUND: the previous samples stay */
    ns = samples_from_readings(i, compute, nc, last_ts, samples);
    sr = ns ? 0 : 6 /* NVML_ERROR_NOT_FOUND */;
  } else {
    sr = n->proc_util(d, samples, &ns, last_ts);
    if (ns > VGPU_MAX_PIDS) ns = VGPU_MAX_PIDS;
    if (g_source == SRC_MIXED) {
      static vgpu_util_sample_t own[VGPU_MAX_PIDS];
      unsigned no = samples_from_readings(i, compute, nc, last_ts, own);
      if (sr != 0) ns = 0;
      for (unsigned a = 0; a < no; a++) { /* a tenant's reading of itself wins over NVML's sample of it */
        unsigned b = 0;
        while (b < ns && samples[b].pid != own[a].pid) b++;
        if (b == ns && ns == VGPU_MAX_PIDS) continue;
        if (b == ns) ns++;
        else { own[a].mem = samples[b].mem; own[a].enc = samples[b].enc; own[a].dec = samples[b].dec; } /* NVML's other engines */
        samples[b] = own[a];
      }
      if (no) sr = 0;
    }
  }

  int fd = wlock(path, i);
  if (fd < 0) return -1;
  vgpu_smutil_dev_t *dev = &file->devices[i];
  dev->compute_size = (uint32_t)nc;
  memcpy(dev->compute, compute, (size_t)nc * sizeof compute[0]);
  dev->graphics_size = (uint32_t)ng;
  memcpy(dev->graphics, graphics, (size_t)ng * sizeof graphics[0]);
  dev->last_seen_us = last_ts;
  if (sr == 0) {
    dev->samples_size = ns;
    memcpy(dev->samples, samples, (size_t)ns * sizeof samples[0]);
  }
  unlock(fd, i);
  return 0;
}

int main(int argc, char **argv) {
  const char *path = VGPU_SMUTIL_FILE, *nvml_path = NULL;
  long passes = -1;
  unsigned period_ms = 80;
  int verbose = 0;
  for (int a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "--file") && a + 1 < argc) path = argv[++a];
    else if (!strcmp(argv[a], "--passes") && a + 1 < argc) passes = atol(argv[++a]);
    else if (!strcmp(argv[a], "--period-ms") && a + 1 < argc) period_ms = (unsigned)atoi(argv[++a]);
    else if (!strcmp(argv[a], "--nvml") && a + 1 < argc) nvml_path = argv[++a];
    else if (!strcmp(argv[a], "--verbose")) verbose = 1;
    else if (!strcmp(argv[a], "--readings-dir") && a + 1 < argc) g_readings_dir = argv[++a];
    else if (!strcmp(argv[a], "--source") && a + 1 < argc && (!strcmp(argv[a + 1], "nvml") || !strcmp(argv[a + 1], "device") || !strcmp(argv[a + 1], "mixed"))) {
      a++;
      g_source = !strcmp(argv[a], "device") ? SRC_DEVICE : !strcmp(argv[a], "mixed") ? SRC_MIXED : SRC_NVML;
    } else {
      fprintf(stderr, "usage: vgpu-smwatcher [--file PATH] [--passes N] [--period-ms 80] [--nvml PATH] [--verbose]\n"
                      "                      [--source nvml|device|mixed] [--readings-dir DIR]\n");
      return 2;
    }
  }
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  nvml_api n;
  if (load_nvml(nvml_path, &n) != 0) return 1;
  vgpu_smutil_t *file = map_file(path);
  if (!file) return 1;
  if (n.init() != 0) { fprintf(stderr, "vgpu-smwatcher: nvmlInit failed\n"); return 1; }
  unsigned count = 0;
  if (n.count(&count) != 0 || count == 0) { fprintf(stderr, "vgpu-smwatcher: no NVML devices\n"); return 1; }
  if (count > VGPU_MAX_DEVICES) count = VGPU_MAX_DEVICES;
  nvml_dev devs[VGPU_MAX_DEVICES];
  for (unsigned i = 0; i < count; i++)
    if (n.by_index(i, &devs[i]) != 0) { fprintf(stderr, "vgpu-smwatcher: no handle for device %u\n", i); return 1; }
  if (verbose) fprintf(stderr, "vgpu-smwatcher: %u device(s) -> %s every %u ms\n", count, path, period_ms);

  /* one pass = all devices; per-device spacing period/count like the Go batches */
  struct timespec gap = {0, (long)((unsigned long long)period_ms * 1000000ull / count)};
  if (gap.tv_nsec >= 1000000000L) { gap.tv_sec = gap.tv_nsec / 1000000000L; gap.tv_nsec %= 1000000000L; }
  int rc = 0;
  for (long pass = 0; !g_stop && (passes < 0 || pass < passes); pass++) {
    for (unsigned i = 0; i < count && !g_stop; i++) {
      if (publish_device(&n, file, path, (int)i, devs[i]) != 0) { rc = 1; g_stop = 1; break; }
      if (passes < 0 || pass + 1 < passes || i + 1 < count) nanosleep(&gap, NULL);
    }
  }
  if (n.shutdown) n.shutdown();
  munmap(file, sizeof *file);
  return rc;
}
