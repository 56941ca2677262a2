/*
 * redirect.c - LD_PRELOAD path-prefix rewriter (TEST INFRASTRUCTURE).
 *
 * The reference compiles its paths in (/etc/vgpu-manager/..., /tmp/.vgpu_lock, /tmp/.vmem_node,
 * reference library/include/hook.h:43-104) and the drop-in library keeps the same paths.  To run
 * either library - and several "containers" side by side - without touching the real /etc or
 * /tmp, the harness preloads this shim *before* the library under test:
 *
 *   VGPU_REDIRECT="/etc/vgpu-manager=/sandbox/etc:/tmp/.vgpu_lock=/sandbox/lock:..."
 *
 * Every path handed to open/fopen/access/mkdir/stat/unlink that starts with a listed prefix is
 * rewritten.  Both libraries see the identical shim, so differential results are unaffected.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#define MAX_RULES 16
static struct { char from[256]; char to[512]; size_t flen; } g_rules[MAX_RULES];
static int g_nrules = -1;

static void load_rules(void) {
  if (g_nrules >= 0) return;
  int n = 0;
  const char *e = getenv("VGPU_REDIRECT");
  if (e) {
    char *dup = strdup(e), *save = NULL;
    for (char *t = strtok_r(dup, ":", &save); t && n < MAX_RULES; t = strtok_r(NULL, ":", &save)) {
      char *eq = strchr(t, '=');
      if (!eq) continue;
      *eq = 0;
      snprintf(g_rules[n].from, sizeof(g_rules[n].from), "%s", t);
      snprintf(g_rules[n].to, sizeof(g_rules[n].to), "%s", eq + 1);
      g_rules[n].flen = strlen(g_rules[n].from);
      n++;
    }
    free(dup);
  }
  g_nrules = n;
}

static const char *rw(const char *path, char *buf, size_t cap) {
  if (!path) return path;
  load_rules();
  for (int i = 0; i < g_nrules; i++) {
    size_t l = g_rules[i].flen;
    if (strncmp(path, g_rules[i].from, l) == 0 && (path[l] == 0 || path[l] == '/')) {
      snprintf(buf, cap, "%s%s", g_rules[i].to, path + l);
      return buf;
    }
  }
  return path;
}

/* dlvsym, not dlsym: the libraries under test interpose `dlsym` itself and would answer an
 * RTLD_NEXT query relative to themselves */
static void *next_sym(const char *name) {
  static const char *vers[] = {"GLIBC_2.2.5", "GLIBC_2.33", "GLIBC_2.34", "GLIBC_2.17", NULL};
  for (int i = 0; vers[i]; i++) {
    void *p = dlvsym(RTLD_NEXT, name, vers[i]);
    if (p) return p;
  }
  return NULL;
}
#define REAL(name) \
  static __typeof__(name) *real_fn = NULL; \
  if (!real_fn) real_fn = (__typeof__(name) *)next_sym(#name)

int open(const char *path, int flags, ...) {
  REAL(open);
  char b[1024];
  mode_t m = 0;
  if (flags & (O_CREAT | O_TMPFILE)) { va_list ap; va_start(ap, flags); m = va_arg(ap, mode_t); va_end(ap); }
  return real_fn(rw(path, b, sizeof b), flags, m);
}
int open64(const char *path, int flags, ...) {
  REAL(open64);
  char b[1024];
  mode_t m = 0;
  if (flags & (O_CREAT | O_TMPFILE)) { va_list ap; va_start(ap, flags); m = va_arg(ap, mode_t); va_end(ap); }
  return real_fn(rw(path, b, sizeof b), flags, m);
}
FILE *fopen(const char *path, const char *mode) {
  REAL(fopen);
  char b[1024];
  return real_fn(rw(path, b, sizeof b), mode);
}
FILE *fopen64(const char *path, const char *mode) {
  REAL(fopen64);
  char b[1024];
  return real_fn(rw(path, b, sizeof b), mode);
}
int access(const char *path, int amode) {
  REAL(access);
  char b[1024];
  return real_fn(rw(path, b, sizeof b), amode);
}
int mkdir(const char *path, mode_t mode) {
  REAL(mkdir);
  char b[1024];
  return real_fn(rw(path, b, sizeof b), mode);
}
int unlink(const char *path) {
  REAL(unlink);
  char b[1024];
  return real_fn(rw(path, b, sizeof b));
}
int stat(const char *path, struct stat *st) {
  REAL(stat);
  char b[1024];
  return real_fn(rw(path, b, sizeof b), st);
}
int execl(const char *path, const char *arg, ...) {
  /* register.c:14 fork/execs the client binary with at most a handful of args */
  char b[1024];
  const char *argv[16];
  int n = 0;
  argv[n++] = arg;
  va_list ap;
  va_start(ap, arg);
  while (n < 15 && argv[n - 1]) argv[n++] = va_arg(ap, const char *);
  va_end(ap);
  argv[n < 16 ? n : 15] = NULL;
  return execv(rw(path, b, sizeof b), (char *const *)argv);
}
