/*
 * scenario.c - a scripted CUDA/NVML tenant used by the differential and GPU tests.
 * TEST INFRASTRUCTURE.
 *
 * It binds the driver the way cudart does - dlopen("libcuda.so.1") + dlsym(), or with
 * `--gpa` through cuGetProcAddress - so an LD_PRELOADed interception library (the reference's
 * or ours) sees every call.  Commands come from stdin, one reply line each; replies contain
 * return codes and sizes only (never pointer values), so the transcripts of two libraries
 * driven by the same script can be compared byte for byte.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>
#include <execinfo.h>
#include <signal.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef struct { size_t W, H; int fmt; unsigned ch; } arr2_t;
typedef struct { size_t W, H, D; int fmt; unsigned ch, flags; } arr3_t;
typedef struct { int type, handle; struct { int type, id; } loc; void *w32; unsigned char flags[8]; } prop_t;
typedef struct { unsigned long long total, free, used; } nvmem_t;
typedef struct { unsigned version; unsigned long long total, reserved, free, used; } nvmem2_t;

static void *h_cuda, *h_nvml;
static int use_gpa;
static CUresult (*p_gpa)(const char *, void **, int, unsigned long long, void *);

static void *sym(const char *name) {
  void *p = NULL;
  if (!strncmp(name, "nvml", 4)) return dlsym(h_nvml, name);
  if (use_gpa && p_gpa) {
    /* cudart asks for base names and lets the driver pick the version */
    char base[96];
    snprintf(base, sizeof base, "%s", name);
    char *v = strstr(base, "_v2");
    if (v) *v = 0;
    if (p_gpa(base, &p, 12090, 0, NULL) == 0 && p) return p;
  }
  return dlsym(h_cuda, name);
}

#define MAXH 4096
static CUdeviceptr g_ptr[MAXH];
static int g_kind[MAXH]; /* 1 linear, 2 array, 3 mipmap, 4 vmm */
static int g_np;
static void *g_exec;

/* a crash inside a preloaded library should leave a trace in the test log: module+offset frames for addr2line */
static void on_crash(int sig) {
  void *frames[48];
  int n = backtrace(frames, 48);
  static const char msg[] = "scenario: fatal signal, backtrace:\n";
  if (write(2, msg, sizeof msg - 1) < 0) _exit(128 + sig);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}

int main(int argc, char **argv) {
  signal(SIGSEGV, on_crash);
  signal(SIGBUS, on_crash);
  signal(SIGABRT, on_crash);
  setvbuf(stdout, NULL, _IOLBF, 0); /* keep the transcript up to the crash */
  for (int i = 1; i < argc; i++)
    if (!strcmp(argv[i], "--gpa")) use_gpa = 1;
  h_cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  h_nvml = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h_cuda || !h_nvml) { fprintf(stderr, "scenario: cannot load driver libs: %s\n", dlerror()); return 2; }
  if (use_gpa) {
    p_gpa = (CUresult(*)(const char *, void **, int, unsigned long long, void *))dlsym(h_cuda, "cuGetProcAddress_v2");
    /* self-lookup, like cudart does: must come back as the interposer's own entry */
    void *again = NULL;
    if (p_gpa && p_gpa("cuGetProcAddress", &again, 12090, 0, NULL) == 0 && again)
      p_gpa = (CUresult(*)(const char *, void **, int, unsigned long long, void *))again;
  }
  setvbuf(stdout, NULL, _IOLBF, 0);
  void *nvdev = NULL;
  void *ctx = NULL;
  int dev = 0;
  char line[4096];
  while (fgets(line, sizeof line, stdin)) {
    char cmd[64] = {0};
    unsigned long long a = 0, b = 0, c = 0, d = 0, e = 0;
    int n = sscanf(line, "%63s %llu %llu %llu %llu %llu", cmd, &a, &b, &c, &d, &e);
    if (n < 1 || cmd[0] == '#') continue;
    if (!strcmp(cmd, "init")) {
      CUresult (*f_init)(unsigned) = sym("cuInit");
      CUresult (*f_get)(int *, int) = sym("cuDeviceGet");
      CUresult (*f_ret)(void **, int) = sym("cuDevicePrimaryCtxRetain");
      CUresult (*f_set)(void *) = sym("cuCtxSetCurrent");
      int (*n_init)(void) = sym("nvmlInit_v2");
      int (*n_h)(unsigned, void **) = sym("nvmlDeviceGetHandleByIndex_v2");
      CUresult r1 = f_init(0);
      CUresult r2 = f_get(&dev, (int)a);
      CUresult r3 = f_ret(&ctx, dev);
      CUresult r4 = f_set(ctx);
      int r5 = n_init();
      int r6 = n_h((unsigned)a, &nvdev);
      printf("init %d %d %d %d %d %d\n", r1, r2, r3, r4, r5, r6);
    } else if (!strcmp(cmd, "drvver")) { /* allowed before cuInit; the hook loads config + device map (cuda_hook.c:1162) */
      CUresult (*f)(int *) = sym("cuDriverGetVersion");
      int v = 0;
      CUresult r = f ? f(&v) : 500;
      printf("drvver -> %d %d\n", r, v);
    } else if (!strcmp(cmd, "dev")) { /* switch this thread to device <a> (its primary context) */
      CUresult (*f_get)(int *, int) = sym("cuDeviceGet");
      CUresult (*f_ret)(void **, int) = sym("cuDevicePrimaryCtxRetain");
      CUresult (*f_set)(void *) = sym("cuCtxSetCurrent");
      int (*n_h)(unsigned, void **) = sym("nvmlDeviceGetHandleByIndex_v2");
      CUresult r2 = f_get(&dev, (int)a);
      CUresult r3 = f_ret(&ctx, dev);
      CUresult r4 = f_set(ctx);
      int r6 = n_h((unsigned)a, &nvdev);
      printf("dev %llu -> %d %d %d %d\n", a, r2, r3, r4, r6);
    } else if (!strcmp(cmd, "reset")) { /* cudaDeviceReset(): primary context destroyed and re-created */
      CUresult (*f_reset)(int) = sym("cuDevicePrimaryCtxReset_v2");
      CUresult (*f_ret)(void **, int) = sym("cuDevicePrimaryCtxRetain");
      CUresult (*f_set)(void *) = sym("cuCtxSetCurrent");
      CUresult r1 = f_reset ? f_reset(dev) : 500;
      g_np = 0; /* every allocation died with the context */
      g_exec = NULL;
      CUresult r2 = f_ret(&ctx, dev);
      CUresult r3 = f_set(ctx);
      printf("reset %d %d %d\n", r1, r2, r3);
    } else if (!strcmp(cmd, "nvmlinit")) {
      /* NVML-only client (nvidia-smi style): no cuInit, no CUDA context */
      int (*n_init)(void) = sym("nvmlInit_v2");
      int (*n_h)(unsigned, void **) = sym("nvmlDeviceGetHandleByIndex_v2");
      int r5 = n_init();
      int r6 = n_h((unsigned)a, &nvdev);
      printf("nvmlinit %d %d\n", r5, r6);
    } else if (!strcmp(cmd, "alloc")) {
      CUresult (*f)(CUdeviceptr *, size_t) = sym("cuMemAlloc_v2");
      CUdeviceptr p = 0;
      CUresult r = f(&p, (size_t)a);
      if (r == 0) { g_ptr[g_np] = p; g_kind[g_np] = 1; g_np++; }
      printf("alloc %llu -> %d h%d\n", a, r, r == 0 ? g_np - 1 : -1);
    } else if (!strcmp(cmd, "managed")) {
      CUresult (*f)(CUdeviceptr *, size_t, unsigned) = sym("cuMemAllocManaged");
      CUdeviceptr p = 0;
      CUresult r = f(&p, (size_t)a, (unsigned)b);
      if (r == 0) { g_ptr[g_np] = p; g_kind[g_np] = 1; g_np++; }
      printf("managed %llu %llu -> %d h%d\n", a, b, r, r == 0 ? g_np - 1 : -1);
    } else if (!strcmp(cmd, "pitch")) {
      CUresult (*f)(CUdeviceptr *, size_t *, size_t, size_t, unsigned) = sym("cuMemAllocPitch_v2");
      CUdeviceptr p = 0;
      size_t pitch = 0;
      CUresult r = f(&p, &pitch, (size_t)a, (size_t)b, (unsigned)c);
      if (r == 0) { g_ptr[g_np] = p; g_kind[g_np] = 1; g_np++; }
      printf("pitch %llu %llu %llu -> %d pitch %zu h%d\n", a, b, c, r, r == 0 ? pitch : 0, r == 0 ? g_np - 1 : -1);
    } else if (!strcmp(cmd, "allocasync")) {
      CUresult (*f)(CUdeviceptr *, size_t, void *) = sym("cuMemAllocAsync");
      CUdeviceptr p = 0;
      CUresult r = f(&p, (size_t)a, NULL);
      if (r == 0) { g_ptr[g_np] = p; g_kind[g_np] = 1; g_np++; }
      printf("allocasync %llu -> %d h%d\n", a, r, r == 0 ? g_np - 1 : -1);
    } else if (!strcmp(cmd, "pool")) {
      CUresult (*f)(CUdeviceptr *, size_t, void *, void *) = sym("cuMemAllocFromPoolAsync");
      CUdeviceptr p = 0;
      CUresult r = f(&p, (size_t)a, NULL, NULL);
      if (r == 0) { g_ptr[g_np] = p; g_kind[g_np] = 1; g_np++; }
      printf("pool %llu -> %d h%d\n", a, r, r == 0 ? g_np - 1 : -1);
    } else if (!strcmp(cmd, "create")) {
      CUresult (*f)(unsigned long long *, size_t, const prop_t *, unsigned long long) = sym("cuMemCreate");
      prop_t pr;
      memset(&pr, 0, sizeof pr);
      pr.type = 1; pr.loc.type = 1; pr.loc.id = dev;
      unsigned long long h = 0;
      CUresult r = f(&h, (size_t)a, &pr, 0);
      if (r == 0) { g_ptr[g_np] = h; g_kind[g_np] = 4; g_np++; }
      printf("create %llu -> %d h%d\n", a, r, r == 0 ? g_np - 1 : -1);
    } else if (!strcmp(cmd, "array")) {
      CUresult (*f)(void **, const arr2_t *) = sym("cuArrayCreate_v2");
      arr2_t ds = {(size_t)a, (size_t)b, (int)c, (unsigned)d};
      void *h = NULL;
      CUresult r = f(&h, &ds);
      if (r == 0) { g_ptr[g_np] = (CUdeviceptr)(uintptr_t)h; g_kind[g_np] = 2; g_np++; }
      printf("array %llu %llu %llu %llu -> %d h%d\n", a, b, c, d, r, r == 0 ? g_np - 1 : -1);
    } else if (!strcmp(cmd, "array3d") || !strcmp(cmd, "mipmap")) {
      arr3_t ds = {(size_t)a, (size_t)b, (size_t)c, (int)d, (unsigned)e, 0};
      void *h = NULL;
      CUresult r;
      if (cmd[0] == 'a') {
        CUresult (*f)(void **, const arr3_t *) = sym("cuArray3DCreate_v2");
        r = f(&h, &ds);
      } else {
        CUresult (*f)(void **, const arr3_t *, unsigned) = sym("cuMipmappedArrayCreate");
        r = f(&h, &ds, 1);
      }
      if (r == 0) { g_ptr[g_np] = (CUdeviceptr)(uintptr_t)h; g_kind[g_np] = cmd[0] == 'a' ? 2 : 3; g_np++; }
      printf("%s %llu %llu %llu %llu %llu -> %d h%d\n", cmd, a, b, c, d, e, r, r == 0 ? g_np - 1 : -1);
    } else if (!strcmp(cmd, "free") || !strcmp(cmd, "freeasync")) {
      CUresult r = 1;
      if ((int)a < g_np && g_ptr[a]) {
        if (g_kind[a] == 1) {
          if (cmd[4]) { CUresult (*f)(CUdeviceptr, void *) = sym("cuMemFreeAsync"); r = f(g_ptr[a], NULL); }
          else { CUresult (*f)(CUdeviceptr) = sym("cuMemFree_v2"); r = f(g_ptr[a]); }
        } else if (g_kind[a] == 2) { CUresult (*f)(void *) = sym("cuArrayDestroy"); r = f((void *)(uintptr_t)g_ptr[a]); }
        else if (g_kind[a] == 3) { CUresult (*f)(void *) = sym("cuMipmappedArrayDestroy"); r = f((void *)(uintptr_t)g_ptr[a]); }
        else { CUresult (*f)(unsigned long long) = sym("cuMemRelease"); r = f(g_ptr[a]); }
        if (r == 0) g_ptr[a] = 0;
      }
      printf("%s h%llu -> %d\n", cmd, a, r);
    } else if (!strcmp(cmd, "dirty")) {
      /* write a pattern over allocation h<a> (first and last byte are what the stub checks) */
      CUresult (*f)(CUdeviceptr, unsigned char, size_t) = sym("cuMemsetD8_v2");
      CUresult r = ((int)a < g_np && g_ptr[a]) ? f(g_ptr[a], 0xA5, (size_t)b) : 1;
      printf("dirty h%llu -> %d\n", a, r);
    } else if (!strcmp(cmd, "fill")) { /* fill h<a> with <b> bytes of value <c> */
      CUresult (*f)(CUdeviceptr, unsigned char, size_t) = sym("cuMemsetD8_v2");
      CUresult (*sy)(void) = sym("cuCtxSynchronize");
      CUresult r = ((int)a < g_np && g_ptr[a]) ? f(g_ptr[a], (unsigned char)c, (size_t)b) : 1;
      if (r == 0 && sy) r = sy();
      printf("fill h%llu -> %d\n", a, r);
    } else if (!strcmp(cmd, "check")) { /* h<a>: do bytes [0,64), the middle and the last 64 of <b> still hold value <c>? */
      CUresult (*f)(void *, CUdeviceptr, size_t) = sym("cuMemcpyDtoH_v2");
      unsigned char buf[64];
      int ok = (int)a < g_np && g_ptr[a] && b >= 64;
      unsigned long long offs[3] = {0, (b / 2) & ~63ull, b - 64};
      CUresult r = 0;
      for (int k = 0; ok && k < 3; k++) {
        r = f(buf, g_ptr[a] + offs[k], 64);
        if (r) { ok = 0; break; }
        for (int i = 0; i < 64; i++) ok &= buf[i] == (unsigned char)c;
      }
      printf("check h%llu -> %d %s\n", a, r, ok ? "intact" : "CORRUPT");
    } else if (!strcmp(cmd, "slabstats")) { /* B200 library only: what the slab mode moved */
      unsigned long long (*metric)(int, int) = dlsym(RTLD_DEFAULT, "vgpu_b200_metric");
      if (metric)
        printf("slabstats allocs %llu demotions %llu spill_bytes %llu promote_bytes %llu scrubbed_bytes %llu spill_ns %llu scrub_ns %llu promote_ns %llu\n",
               metric((int)a, 16), metric((int)a, 17), metric((int)a, 11), metric((int)a, 14), metric((int)a, 8), metric((int)a, 12),
               metric((int)a, 13), metric((int)a, 15));
      else printf("slabstats none\n");
    } else if (!strcmp(cmd, "meminfo")) {
      CUresult (*f)(size_t *, size_t *) = sym("cuMemGetInfo_v2");
      size_t fr = 0, tot = 0;
      CUresult r = f(&fr, &tot);
      printf("meminfo -> %d free %zu total %zu\n", r, fr, tot);
    } else if (!strcmp(cmd, "totalmem")) {
      CUresult (*f)(size_t *, int) = sym("cuDeviceTotalMem_v2");
      size_t tot = 0;
      CUresult r = f(&tot, dev);
      printf("totalmem -> %d %zu\n", r, tot);
    } else if (!strcmp(cmd, "nvmlinfo")) {
      int (*f)(void *, nvmem_t *) = sym("nvmlDeviceGetMemoryInfo");
      nvmem_t m = {0, 0, 0};
      int r = f(nvdev, &m);
      printf("nvmlinfo -> %d total %llu free %llu used %llu\n", r, m.total, m.free, m.used);
    } else if (!strcmp(cmd, "nvmlinfo2")) {
      int (*f)(void *, nvmem2_t *) = sym("nvmlDeviceGetMemoryInfo_v2");
      nvmem2_t m;
      memset(&m, 0, sizeof m);
      m.version = (unsigned)(sizeof m | (2u << 24));
      int r = f(nvdev, &m);
      printf("nvmlinfo2 -> %d total %llu reserved %llu free %llu used %llu\n", r, m.total, m.reserved, m.free, m.used);
    } else if (!strcmp(cmd, "setmode")) {
      int (*f)(void *, int) = sym("nvmlDeviceSetComputeMode");
      printf("setmode -> %d\n", f(nvdev, (int)a));
    } else if (!strcmp(cmd, "persistence")) {
      int (*f)(void *, int *) = sym("nvmlDeviceGetPersistenceMode");
      int m = -1;
      int r = f(nvdev, &m);
      printf("persistence -> %d mode %d\n", r, m);
    } else if (!strcmp(cmd, "launch")) {
      CUresult (*f)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **) =
          sym("cuLaunchKernel");
      unsigned long long okc = 0;
      struct timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      for (unsigned long long i = 0; i < a; i++) okc += f(NULL, (unsigned)b, (unsigned)c, (unsigned)d, 1, 1, 1, 0, NULL, NULL, NULL) == 0;
      clock_gettime(CLOCK_MONOTONIC, &t1);
      double sec = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
      printf("launch %llu -> ok %llu\n", a, okc);
      fprintf(stderr, "launch_rate %.0f per_s\n", sec > 0 ? a / sec : 0.0);
    } else if (!strcmp(cmd, "launchvia")) {
      /* launchvia <kind> <n> <gx> <gy>: the same launch through each of the entry points the reference hooks
       * (cuda_hook.c:1810-2002).  0 cuLaunchKernel, 1 _ptsz, 2 cuLaunchKernelEx, 3 _ptsz, 4 cuLaunchCooperativeKernel,
       * 5 _ptsz, 6 cuLaunchGrid, 7 cuLaunchGridAsync, 8 cuFuncSetBlockShape + cuLaunch */
      typedef CUresult (*k_fn)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **);
      typedef CUresult (*c_fn)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **);
      struct { unsigned gx, gy, gz, bx, by, bz, smem; void *stream; void *attrs; unsigned nattrs; } cfg =
          {(unsigned)c, (unsigned)d, 1, 1, 1, 1, 0, NULL, NULL, 0};
      static const char *names[] = {"cuLaunchKernel", "cuLaunchKernel_ptsz", "cuLaunchKernelEx", "cuLaunchKernelEx_ptsz",
                                    "cuLaunchCooperativeKernel", "cuLaunchCooperativeKernel_ptsz", "cuLaunchGrid",
                                    "cuLaunchGridAsync", "cuLaunch"};
      unsigned long long okc = 0;
      if (a > 8) { printf("launchvia %llu -> unknown kind\n", a); continue; }
      void *fn = sym(names[a]);
      if (a == 8) {
        CUresult (*shape)(void *, int, int, int) = sym("cuFuncSetBlockShape");
        printf("blockshape -> %d\n", shape ? shape(NULL, 4, 2, 1) : -1);
      }
      for (unsigned long long i = 0; fn && i < b; i++) {
        CUresult r;
        switch (a) {
        case 0: case 1: r = ((k_fn)fn)(NULL, (unsigned)c, (unsigned)d, 1, 1, 1, 1, 0, NULL, NULL, NULL); break;
        case 2: case 3: r = ((CUresult(*)(const void *, void *, void **, void **))fn)(&cfg, NULL, NULL, NULL); break;
        case 4: case 5: r = ((c_fn)fn)(NULL, (unsigned)c, (unsigned)d, 1, 1, 1, 1, 0, NULL, NULL); break;
        case 6: r = ((CUresult(*)(void *, int, int))fn)(NULL, (int)c, (int)d); break;
        case 7: r = ((CUresult(*)(void *, int, int, void *))fn)(NULL, (int)c, (int)d, NULL); break;
        default: r = ((CUresult(*)(void *))fn)(NULL); break;
        }
        okc += r == 0;
      }
      printf("launchvia %s %llu -> ok %llu\n", names[a], b, okc);
    } else if (!strcmp(cmd, "ledger")) {
      /* raw bytes of this GPU's record in vmem_node.config; pids are normalised to rank order */
      const char *path = getenv("SCENARIO_LEDGER");
      FILE *fp = path ? fopen(path, "rb") : NULL;
      if (!fp) { printf("ledger -> absent\n"); continue; }
      static unsigned char rec[16392];
      fseek(fp, (long)(a * 16392), SEEK_SET);
      size_t got = fread(rec, 1, sizeof rec, fp);
      fclose(fp);
      unsigned sz = 0;
      if (got == sizeof rec) memcpy(&sz, rec + 16384, 4);
      printf("ledger -> size %u", sz);
      for (unsigned i = 0; i < sz && i < 8; i++) {
        int pid; unsigned long long used;
        memcpy(&pid, rec + i * 16, 4);
        memcpy(&used, rec + i * 16 + 8, 8);
        printf(" [%s %llu]", pid == getpid() ? "self" : "other", used);
      }
      printf("\n");
    } else if (!strcmp(cmd, "forkchild")) {
      /* fork; the child allocates <a> bytes, reads meminfo, launches <b> kernels, exits */
      fflush(stdout);
      pid_t pid = fork();
      if (pid == 0) {
        CUresult (*fa)(CUdeviceptr *, size_t) = sym("cuMemAlloc_v2");
        CUresult (*fm)(size_t *, size_t *) = sym("cuMemGetInfo_v2");
        CUresult (*fl)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **) =
            sym("cuLaunchKernel");
        CUdeviceptr p = 0;
        size_t fr = 0, tot = 0;
        CUresult r1 = fa(&p, (size_t)a);
        CUresult r2 = fm(&fr, &tot);
        unsigned long long okc = 0;
        for (unsigned long long i = 0; i < b; i++) okc += fl(NULL, 1, 1, 1, 1, 1, 1, 0, NULL, NULL, NULL) == 0;
        printf("child alloc %d meminfo %d total %zu launches %llu\n", r1, r2, tot, okc);
        fflush(stdout);
        _exit(0);
      }
      int st = 0;
      waitpid(pid, &st, 0);
      printf("forkchild -> exit %d\n", WIFEXITED(st) ? WEXITSTATUS(st) : -1);
    } else if (!strcmp(cmd, "graph")) { /* graph <kernel nodes> <gridX>: build + instantiate (replaces the previous one) */
      typedef struct { void *func; unsigned gx, gy, gz, bx, by, bz, smem; void **params; void **extra; } knp_t;
      CUresult (*create)(void **, unsigned) = sym("cuGraphCreate");
      CUresult (*add)(void **, void *, const void **, size_t, const knp_t *) = sym("cuGraphAddKernelNode");
      CUresult (*inst)(void **, void *, unsigned long long) = sym("cuGraphInstantiateWithFlags");
      CUresult (*gdestroy)(void *) = sym("cuGraphDestroy");
      CUresult (*edestroy)(void *) = sym("cuGraphExecDestroy");
      CUresult r = 500;
      if (create && add && inst) {
        if (g_exec && edestroy) edestroy(g_exec);
        g_exec = NULL;
        void *g = NULL;
        r = create(&g, 0);
        for (unsigned long long i = 0; r == 0 && i < a; i++) {
          knp_t p = {NULL, (unsigned)b, 1, 1, 1, 1, 1, 0, NULL, NULL};
          void *node = NULL;
          r = add(&node, g, NULL, 0, &p);
        }
        if (r == 0) r = inst(&g_exec, g, 0);
        if (g && gdestroy) gdestroy(g);
      }
      printf("graph %llu x %llu -> %d\n", a, b, r);
    } else if (!strcmp(cmd, "graphlaunch")) {
      CUresult (*gl)(void *, void *) = sym("cuGraphLaunch");
      unsigned long long okc = 0;
      for (unsigned long long i = 0; gl && g_exec && i < a; i++) okc += gl(g_exec, NULL) == 0;
      printf("graphlaunch %llu -> ok %llu\n", a, okc);
    } else if (!strcmp(cmd, "limstate")) { /* B200 library only: tokens charged so far */
      struct { long long granted, consumed, bucket, share; int v[10]; unsigned long long steps; } ls;
      int (*lstate)(void *) = dlsym(RTLD_DEFAULT, "vgpu_b200_limiter_state");
      memset(&ls, 0, sizeof ls);
      if (lstate && lstate(&ls) == 0) printf("limstate consumed %lld\n", ls.consumed);
      else printf("limstate none\n");
    } else if (!strcmp(cmd, "metrics")) { /* B200 library only */
      unsigned long long (*metric)(int, int) = dlsym(RTLD_DEFAULT, "vgpu_b200_metric");
      struct { long long granted, consumed, bucket, share; int v[10]; unsigned long long steps; } ls;
      int (*lstate)(void *) = dlsym(RTLD_DEFAULT, "vgpu_b200_limiter_state");
      memset(&ls, 0, sizeof ls);
      if (metric && lstate && lstate(&ls) == 0)
        printf("metrics sampler_launches %llu sampler_skipped %llu control_steps %llu\n", metric((int)a, 7), metric((int)a, 10), ls.steps);
      else printf("metrics none\n");
    } else if (!strcmp(cmd, "sleepms")) {
      struct timespec ts = {(time_t)(a / 1000), (long)(a % 1000) * 1000000L};
      nanosleep(&ts, NULL);
      printf("sleepms %llu\n", a);
    } else {
      printf("unknown %s\n", cmd);
    }
  }
  return 0;
}
