/*
 * vgpu_oracle.h - CPU restatement of the reference's two enforcement algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vgpu_manager_b200/ may include, link or dlopen
 * this.  Allowed users: tests/, tests/stub (the fake GPU used for CPU-only plumbing tests),
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg - always as the checker.
 *
 * Parity status: PINNED DIFFERENTIALLY.  The reference ships no golden vectors for this
 * path, so every function here is pinned against the reference's own code
 * executed in this container: oracle/_ref/libvgpu-control.so (the reference library built
 * from /root/reference/library/src by oracle/Makefile) and oracle/_ref/ref_cosim (ref_cosim.c
 * #includes the reference's cuda_hook.c, so its `static` arithmetic and its watcher thread run
 * unmodified, the latter on a virtual clock).  tests/golden/ holds the vectors generated from those.
 *
 * Each function cites the reference lines it follows.
 */
#ifndef VGPU_ORACLE_H
#define VGPU_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/vgpu_contract.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- compute-share limiter ---------------- */

typedef struct {
  int32_t sm_num;            /* CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT            */
  int32_t max_thread_per_sm; /* CU_DEVICE_ATTRIBUTE_MAX_THREADS_PER_MULTIPROCESSOR  */
  int64_t total_cores;       /* thr * sm * 32  (cuda_hook.c:534)                    */
} orc_gpu_t;

void orc_gpu_init(orc_gpu_t *g, int sm_num, int max_thread_per_sm);

/* cuda_hook.c:332-367 */
int64_t orc_delta(const orc_gpu_t *g, int up_limit, int user_current, int64_t share);
/* cuda_hook.c:292-306: returns the new bucket value */
int64_t orc_change_token(const orc_gpu_t *g, int64_t bucket, int64_t delta);
/* cuda_hook.c:308-330 + the unsigned product / int cast at the call sites (:1822, :1842).
 * Returns 1 and updates *bucket if the launch is admitted, 0 if the caller must wait. */
int orc_rate_limiter_try(int64_t *bucket, uint32_t gx, uint32_t gy, uint32_t gz);

typedef struct {
  int32_t user_current;
  int32_t sys_current;
  int32_t valid;
  int32_t sys_process_num;
} orc_util_t;

/* per-device statics of the watcher (cuda_hook.c:369-378) */
typedef struct {
  int64_t share;
  int32_t sys_free;
  int32_t avg_sys_free;
  int32_t i;
  int32_t pre_sys_process_num;
  int32_t up_limit;
  int32_t _pad;
} orc_watcher_t;

/* cuda_hook.c:392-401 */
void orc_watcher_init(orc_watcher_t *w, const vgpu_cfg_dev_t *cfg);
/* one iteration of the loop body cuda_hook.c:413-466 for one device, given the utilisation
 * reading it would have obtained.  Updates *w and *bucket. */
void orc_watcher_step(const orc_gpu_t *g, const vgpu_cfg_dev_t *cfg, orc_watcher_t *w,
                      const orc_util_t *u, int64_t *bucket);

/* cuda_hook.c:1057-1155: fold per-process samples into user/sys utilisation.
 * `primary[i]` / `local[i]` are the two container-membership predicates of sample i
 * (primary = the mode's own test, local = check_device_pid_in_local_container_pid);
 * ignored in HOST mode.  `u->valid` is sticky exactly like the reference. */
void orc_fold_utilization(int mode, const vgpu_util_sample_t *s, uint32_t n, uint64_t checktime,
                          const uint8_t *primary, const uint8_t *local, int have_container_pids,
                          orc_util_t *u);

/* cuda_hook.c:541-564: fills start/end per batch, returns batch count */
int orc_balance_batches(int device_count, int sm_watcher, int *start, int *end);

/* ---------------- memory cap / oversubscription ---------------- */

enum { ORC_PATH_GPU = 0, ORC_PATH_UVA = 1, ORC_PATH_OOM = 2 };

/* cuda_hook.c:735-805 (latching semantics of matchX/matchOpenKernel) */
uint64_t orc_accumulate_used(int mode, const vgpu_proc_t *p, uint32_t n, const uint8_t *primary,
                             const uint8_t *local);
/* cuda_hook.c:807-891: compute list + graphics list with pid dedup.  The flag arrays are
 * indexed like the *input* lists. */
uint64_t orc_used_memory(int mode, const vgpu_proc_t *comp, uint32_t nc, const uint8_t *cprim,
                         const uint8_t *cloc, const vgpu_proc_t *gfx, uint32_t ng,
                         const uint8_t *gprim, const uint8_t *gloc);
/* cuda_hook.c:93-116 (after load_limited_memory_view succeeded) */
int orc_memory_path(const vgpu_cfg_dev_t *cfg, uint64_t used, uint64_t vmem, uint64_t request,
                    int allow_uva);
/* nvml_hook.c:58-63 / :89-98 */
void orc_nvml_meminfo(const vgpu_cfg_dev_t *cfg, uint64_t used, uint64_t vmem, uint64_t *total,
                      uint64_t *out_used, uint64_t *out_free);
/* cuda_hook.c:1741-1787 */
void orc_cu_meminfo(const vgpu_cfg_dev_t *cfg, uint64_t used, uint64_t vmem, int real_ok,
                    uint64_t real_total, uint64_t *free_out, uint64_t *total_out);

/* request sizes: cuda_hook.c:138-146,1546-1569 (bits, not bytes!) and :1411-1412 */
uint64_t orc_array_request(int format, uint64_t channels, uint64_t h, uint64_t w);
uint64_t orc_array3d_request(int format, uint64_t channels, uint64_t h, uint64_t w, uint64_t d);
uint64_t orc_pitch_guess(uint64_t width_bytes, uint32_t elem);

/* UVA ledger, loader.c:1824-1922 and :1580-1619 */
int orc_ledger_add(vgpu_vmem_dev_t *d, int pid, uint64_t bytes); /* 0 ok, -1 table full */
void orc_ledger_sub(vgpu_vmem_dev_t *d, int pid, uint64_t bytes);
uint64_t orc_ledger_sum(const vgpu_vmem_dev_t *d);
void orc_ledger_rm_pid(vgpu_vmem_dev_t *d, int pid);
/* loader.c:1580-1602; alive[i] != 0 <=> record i's pid exists and is not a zombie */
void orc_ledger_purge(vgpu_vmem_dev_t *d, int self_pid, const uint8_t *alive);

/* ---------------- env -> vgpu.config ---------------- */

typedef const char *(*orc_getenv_fn)(const char *name, void *ctx);
/* util.c:27-53 */
uint64_t orc_iec_to_bytes(const char *s);
/* loader.c:1927-2052 with util.c:55-213 */
void orc_config_from_env(orc_getenv_fn ge, void *ctx, vgpu_cfg_t *out);

#ifdef __cplusplus
}
#endif
#endif
