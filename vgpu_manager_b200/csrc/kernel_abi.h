/*
 * kernel_abi.h - parameter blocks shared by the host shim (C), the sm_100a kernels (CUDA) and
 * the CPU-only fake GPU used by the plumbing tests (tests/stub).  Plain C, fixed-width types,
 * no padding surprises (static asserts below).
 */
#ifndef VGPU_KERNEL_ABI_H
#define VGPU_KERNEL_ABI_H

#include <stdint.h>
#include "../../include/vgpu_contract.h"

#ifdef __cplusplus
extern "C" {
#endif

/* kernel entry names inside the embedded sm_100a image */
#define VGPU_K_CLEAR "vgpu_clear_kernel"
#define VGPU_K_SPILL "vgpu_spill_copy_kernel"
#define VGPU_K_QUOTA "vgpu_quota_kernel"
#define VGPU_K_SLAB_INSERT "vgpu_slab_insert_kernel"
#define VGPU_K_SLAB_REMOVE "vgpu_slab_remove_kernel"
#define VGPU_K_CONTROLLER "vgpu_controller_kernel"
#define VGPU_K_SAMPLER "vgpu_sampler_kernel"
#define VGPU_K_GATE "vgpu_gate_kernel"
#define VGPU_K_GOVERNOR "vgpu_governor_kernel"
#define VGPU_K_REFILL "vgpu_refill_kernel"

/* ---------------------------------------------------------------- spill copy geometry */
#define VGPU_SPILL_CHUNK 16384u      /* bytes per TMA bulk copy                      */
#define VGPU_SPILL_STAGES 4u         /* shared-memory ring depth per CTA             */
#define VGPU_SPILL_SMEM_BYTES (VGPU_SPILL_CHUNK * VGPU_SPILL_STAGES)
#define VGPU_SPILL_CTAS_PER_SM 1u    /* one 64 KiB ring per SM (best of the r1 sweep)  */

/* ---------------------------------------------------------------- memory quota */

#define VGPU_FLAG_PRIMARY 1u /* pid passes the compatibility mode's own container test */
#define VGPU_FLAG_LOCAL 2u   /* pid passes the open-kernel "local container pid" test   */

enum { VGPU_Q_ALLOC = 0, VGPU_Q_NVML_INFO = 1, VGPU_Q_CU_INFO = 2 };
enum { VGPU_PATH_GPU = 0, VGPU_PATH_UVA = 1, VGPU_PATH_OOM = 2,
       VGPU_PATH_RETRY = 3 /* armed launch gave up (timeout / CTA too small): evaluate again */ };

/* request block: pinned, device-mapped host memory, one per GPU; written by the host under the
 * per-GPU file lock, read by the kernel with coalesced 128-bit loads */
typedef struct {
  uint32_t seq;          /* echoed into the result when the kernel is done           */
  uint32_t kind;         /* VGPU_Q_*                                                 */
  uint32_t mode;         /* compatibility mode                                       */
  uint32_t allow_uva;
  uint32_t memory_oversold;
  uint32_t real_ok;      /* CU_INFO: driver's cuMemGetInfo succeeded                 */
  uint32_t n_compute;
  uint32_t n_graphics;
  uint32_t n_vmem;       /* ledger records of this GPU (0 when vmem_node is off)     */
  uint32_t _pad;
  uint64_t total_memory;
  uint64_t real_memory;
  uint64_t request;
  uint64_t real_total;   /* CU_INFO: driver total                                    */
  uint64_t self_bytes;   /* this library's own device footprint, removed from `used` */
  uint32_t self_pid;
  uint32_t _pad2[3];     /* record arrays start 16-byte aligned: they are read with 128-bit loads */
  vgpu_proc_t compute[VGPU_MAX_PIDS];
  vgpu_proc_t graphics[VGPU_MAX_PIDS];
  vgpu_vmem_rec_t vmem[VGPU_MAX_PIDS];
  uint8_t cflags[VGPU_MAX_PIDS];
  uint8_t gflags[VGPU_MAX_PIDS];
} vgpu_quota_req_t;

typedef struct {
  uint64_t used;     /* container's physical bytes (NVML view)              */
  uint64_t vmem;     /* ledger sum                                          */
  uint64_t total;    /* reported total                                      */
  uint64_t out_used; /* reported used                                       */
  uint64_t out_free; /* reported free                                       */
  uint32_t path;     /* VGPU_PATH_*                                         */
  uint32_t seq_done; /* written last (release, system scope)                */
} vgpu_quota_res_t;
VGPU_STATIC_ASSERT(offsetof(vgpu_quota_req_t, compute) % 16 == 0, qreq_compute_align);
VGPU_STATIC_ASSERT(offsetof(vgpu_quota_req_t, graphics) % 16 == 0, qreq_graphics_align);
VGPU_STATIC_ASSERT(offsetof(vgpu_quota_req_t, vmem) % 16 == 0, qreq_vmem_align);

/* ---------------------------------------------------------------- UVA slab ledger */

#define VGPU_SLAB_SLOTS 16384u /* power of two; 16 B each = 256 KiB of HBM */
typedef struct {
  uint64_t dptr; /* 0 = free, 1 = tombstone */
  uint64_t bytes;
} vgpu_slab_slot_t;

typedef struct {
  uint64_t bytes;    /* remove: size of the record found (0 if none)        */
  uint32_t slot;     /* slot index used / found, 0xffffffff if none         */
  uint32_t seq_done;
} vgpu_slab_res_t;

/* ---------------------------------------------------------------- slab placement table (VGPU_B200_SLAB=1)
 * Device-resident table of the VMM-backed slabs the allocation hooks serve cuMemAlloc from when a
 * device is oversold: which slabs are accounted like the reference's GPU path / UVA path (the
 * accounting is exactly the reference's), and where their backing currently lives (HBM or host).
 * vgpu_vslab_kernel is the allocator's bookkeeping: free-slot scan (PUT), lookup (TAKE) and
 * the spill decision - the coldest slab of a size class in a given placement (SCAN), flipped in
 * the same launch. */
#define VGPU_K_VSLAB "vgpu_vslab_kernel"
#define VGPU_VSLAB_SLOTS 4096u
enum { VGPU_VSLAB_PUT = 0, VGPU_VSLAB_TAKE = 1, VGPU_VSLAB_SCAN = 2 };
#define VGPU_VS_UVA 1u /* accounted on the reference's UVA path: its bytes are in the ledger     */
#define VGPU_VS_DEV 2u /* backing currently in HBM (else: host memory mapped under the same VA)  */
typedef struct {
  uint64_t dptr;  /* 0 = free slot */
  uint64_t bytes; /* what the tenant asked for                  */
  uint64_t size;  /* mapped size (granularity multiple) = size class */
  uint32_t state; /* VGPU_VS_*                                  */
  uint32_t age;   /* allocation order; smallest = coldest       */
} vgpu_vslab_slot_t;
typedef struct {
  uint32_t op;               /* VGPU_VSLAB_*                                                   */
  uint32_t mask, want;       /* SCAN: (state & mask) == want                                   */
  uint32_t set_mask, set_val;/* SCAN: new state bits of the slot found                         */
  uint32_t age;              /* PUT                                                            */
  uint64_t dptr;             /* PUT / TAKE key; SCAN: 0 = any slab of the size class           */
  uint64_t bytes, size;      /* PUT; SCAN: size class                                          */
  uint32_t state, _pad;      /* PUT                                                            */
} vgpu_vslab_req_t;
typedef struct {
  uint64_t dptr, bytes, size;
  uint32_t state, age;
  uint32_t slot;     /* 0xffffffff: nothing found / table full */
  uint32_t seq_done;
} vgpu_vslab_res_t;
VGPU_STATIC_ASSERT(sizeof(vgpu_vslab_slot_t) == 32, vslab_slot);

/* ---------------------------------------------------------------- limiter */

#define VGPU_STREAM_SLOTS 64u
#define VGPU_TICKET_RING 1024u /* per stream slot; power of two */
#define VGPU_MAX_SMS 256u
enum { VGPU_SRC_QUEUE = 0, VGPU_SRC_SM = 1, VGPU_SRC_MAX = 2, VGPU_SRC_NVML = 3 }; /* vgpu_lim_host_t.util_source */
#define VGPU_SAMPLER_PROBE_ONLY 0x7fffffffu /* sampler period_ticks value: probe SMs, leave the queue signal and the controller to the governor */

/* device-resident (HBM) limiter state, one per GPU */
typedef struct {
  /* token bucket: tokens = granted - consumed (consumed lives in the host block) */
  int64_t granted;
  int64_t bucket_last;
  /* controller statics == reference cuda_hook.c:369-378 */
  int64_t share;
  int32_t sys_free;
  int32_t avg_sys_free;
  int32_t ctr_i;
  int32_t pre_sys_process_num;
  int32_t up_limit;
  int32_t valid;
  /* configuration */
  int32_t hard_core, soft_core, core_limit, hard_limit;
  int32_t sm_num, max_thread_per_sm;
  int64_t total_cores;
  /* sampler accumulators for the current control period */
  unsigned long long busy_samples;
  unsigned long long total_samples;
  unsigned long long probe_cycles;      /* sum over SM sub-partitions of measured probe cycles */
  unsigned long long probe_idle_cycles; /* sum of the calibrated idle cost of those probes      */
  unsigned long long probe_count;
  uint32_t cta_done;                    /* ticket for "last CTA runs the controller"            */
  uint32_t period_tick;                 /* sampler launches since the last control step         */
  uint32_t quit_dev;                    /* set by CTA 0 when the host asked the sampler to leave */
  uint32_t _pad_q;
  uint32_t probe_idle[VGPU_MAX_SMS * 4];/* per SM sub-partition calibrated idle probe cycles    */
  uint32_t sm_epoch[VGPU_MAX_SMS];      /* last sampler launch (epoch) that ran on this %smid   */
  /* utilisation history: the controller is fed the mean of the last `util_window` periods, the
   * on-device stand-in for NVML's ~1 s process-utilisation window (cuda_hook.c:972-976) */
  int32_t util_hist[16];
  uint32_t util_hist_pos;
  int32_t blk_sum, blk_n, blk_reading; /* tumbling-block mode: reading = mean of the last full block */
  unsigned long long last_ctl_ns;      /* %globaltimer of the last control step (survives governor restarts) */
  unsigned long long gov_left_ns;      /* %globaltimer at which the previous governor incarnation retired   */
  uint32_t gov_left_busy;              /* ... and whether tenant work was executing at that moment           */
  uint32_t _pad_g;
  /* node-level rebalance: last applied sequence number and the assigned target (0 = none) */
  uint32_t limits_seen;
  int32_t ext_up;
  /* top_result of the NVML-sample reading (cuda_hook.c:376): persists between publications */
  int32_t top_user, top_sys, top_nproc, top_seq;
  /* last results, for metrics and tests */
  int32_t last_user_current;
  int32_t last_sys_current;
  int32_t last_sm_active_pct;
  int32_t last_queue_busy_pct;
  unsigned long long steps;
} vgpu_lim_dev_t;

/* pinned, device-mapped, portable host block: what the launch hook touches */
typedef struct {
  volatile long long consumed;        /* host-owned, fetch_add by the hook                     */
  volatile long long granted_mirror;  /* device-written copy of granted                        */
  volatile uint32_t quit;             /* >0: threads inside a device-wide synchronise; resident kernels leave */
  volatile uint32_t util_source;      /* VGPU_SRC_*: 3 NVML samples (default), 0 queue-busy, 1 sm-active, 2 max of both */
  volatile int32_t ext_sys_current;   /* other tenants' util (host-provided, balance mode)     */
  volatile int32_t ext_sys_process_num;
  volatile int32_t ext_user_override; /* >=0: test hook, use this as user_current              */
  volatile uint32_t util_window;      /* periods averaged into user_current (1..16)            */
  volatile uint32_t util_mode;        /* 0 moving average, 1 tumbling block (NVML-like sampling) */
  volatile uint32_t ctl_state;        /* governor: 0 not resident, 1 resident (or being launched), 2 leaving */
  volatile uint32_t release_pending;  /* watchdog lent tokens: fold release_floor into granted at the next step */
  volatile uint32_t ext_limits_seq;   /* node-level rebalance (rebalance.config): bumped when the two values below change */
  volatile int32_t ext_up_limit;      /* utilisation target the node agent assigned to this tenant (percent)            */
  volatile int32_t ext_soft_core;     /* ceiling it may be raised to; > hard_core switches the device to balance mode    */
  volatile long long release_floor;   /* ticket up to which the watchdog released parked launches */
  volatile unsigned long long launched[VGPU_STREAM_SLOTS]; /* per-slot launch sequence (host)   */
  /* per-slot completion markers, written by cuStreamWriteValue64 right after each launch */
  volatile unsigned long long done[VGPU_STREAM_SLOTS];
  /* ticket (= `consumed` before the launch) of launch `seq` of slot s at [s][seq % RING] */
  volatile long long ticket[VGPU_STREAM_SLOTS][VGPU_TICKET_RING];
  /* mirrors for the host (metrics / tests) */
  volatile int32_t user_current, sys_current, sm_active_pct, queue_busy_pct;
  volatile long long share_mirror, bucket_mirror;
  volatile int32_t up_limit_mirror;
  volatile uint32_t gov_left_busy;    /* the governor retired while tenant work was executing (see device block) */
  volatile unsigned long long steps;
} vgpu_lim_host_t;

/* ---------------------------------------------------------------- utilisation reading (L5)
 * One publication per control period by the tick thread: the raw per-process samples exactly as
 * nvmlDeviceGetProcessUtilization (or sm_util.config) returned them plus the container
 * membership flags of their pids.  vgpu_refill_kernel folds them (reference
 * cuda_hook.c:1044-1159) and runs the controller step (:413-466) in the same launch.
 * Pinned, device-mapped host memory; record arrays 16-byte aligned (128-bit loads). */
enum {
  VGPU_UTIL_NOTHING = 0,   /* process-list query failed: top_result untouched (:946-950)         */
  VGPU_UTIL_NPROC_ONLY = 1,/* list ok, sample query failed (NOT_FOUND between NVML updates):
                              only sys_process_num changes, the previous reading is kept (:984) */
  VGPU_UTIL_SAMPLES = 2    /* both ok: fold the samples                                          */
};
typedef struct {
  uint32_t seq;
  uint32_t status;           /* VGPU_UTIL_*                                        */
  uint32_t mode;             /* compatibility mode                                 */
  uint32_t n_samples;
  int32_t sys_process_num;
  uint32_t have_container_pids; /* client mode: pids.config non-empty (:1073)      */
  uint32_t _pad[2];
  uint64_t checktime_us;     /* samples older than this are skipped (:972-979)     */
  uint64_t _pad2;
  vgpu_util_sample_t samples[VGPU_MAX_PIDS];
  uint8_t flags[VGPU_MAX_PIDS]; /* VGPU_FLAG_* per sample pid                      */
} vgpu_util_req_t;
VGPU_STATIC_ASSERT(offsetof(vgpu_util_req_t, samples) % 16 == 0, ureq_samples_align);

/* ---------------------------------------------------------------- node-level rebalance
 * VGPU_CFG_DIR/rebalance.config, written by a node agent (control-plane side of the read-only
 * config mount, so a tenant cannot raise its own share), read by the tick thread once per control
 * period.  No reference counterpart: the reference's balance policy lives inside
 * each process (cuda_hook.c:430-465). */
#define VGPU_REBALANCE_MAGIC 0x4c424756u /* "VGBL" */
typedef struct {
  uint32_t magic;
  uint32_t seq;       /* bumped by the agent with every change; 0 = nothing assigned */
  int32_t up_limit;   /* utilisation target in percent                               */
  int32_t soft_core;  /* ceiling; <= hard_core leaves the device in hard-limit mode   */
} vgpu_rebalance_rec_t;
typedef struct {
  vgpu_rebalance_rec_t devices[VGPU_MAX_DEVICES];
} vgpu_rebalance_t;
/* VGPU_LOCK_DIR/vgpu_<h>.status, written by a tenant's tick thread while a rebalance.config
 * exists: what the agent folds into its next plan */
typedef struct {
  uint32_t magic, seq;
  int32_t pid, user_current, sys_current, up_limit, hard_core, soft_core;
  int64_t share;
  uint64_t gated, launched, steps;
} vgpu_tenant_status_t;

/* one explicit controller step (also the unit the parity tests drive) */
typedef struct {
  int32_t user_current;
  int32_t sys_current;
  int32_t valid;
  int32_t sys_process_num;
} vgpu_ctrl_in_t;

#ifdef __cplusplus
}
#endif
#endif
