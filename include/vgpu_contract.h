/*
 * vgpu_contract.h - the file / env contract between the (unchanged) Go control plane and
 * the interception library.  Everything here is dictated by the reference and is byte-for-byte
 * what its device-plugin writes and its monitor reads:
 *
 *   resource_data_t  -> vgpu_cfg_t        reference library/include/hook.h:161-189
 *   device_util_t    -> vgpu_smutil_t     reference library/include/hook.h:200-220
 *   device_vmemory_t -> vgpu_vmem_t       reference library/include/hook.h:228-241
 *   paths                                  reference library/include/hook.h:43-104
 *   Go writer of vgpu.config               reference pkg/config/vgpu/vgpu_config.go:46-97
 *   Go reader of vmem_node.config          reference pkg/config/vmem/vmem_config.go:179-205
 *
 * Sizes are pinned with static asserts (1848 / 1311232 / 262272 bytes).
 */
#ifndef VGPU_CONTRACT_H
#define VGPU_CONTRACT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGPU_MAX_DEVICES 16   /* hook.h:116 */
#define VGPU_MAX_PIDS 1024    /* hook.h:121 */
#define VGPU_UUID_LEN 48      /* hook.h:122 */
#define VGPU_NAME_LEN 64      /* hook.h:123 */

/* fixed absolute paths (hook.h:43-104) */
#define VGPU_ROOT_DIR "/etc/vgpu-manager"
#define VGPU_CFG_DIR VGPU_ROOT_DIR "/config"
#define VGPU_CFG_FILE VGPU_CFG_DIR "/vgpu.config"
#define VGPU_PIDS_FILE VGPU_CFG_DIR "/pids.config"
#define VGPU_REBALANCE_FILE VGPU_CFG_DIR "/rebalance.config" /* B200 addition, optional (kernel_abi.h) */
#define VGPU_STATUS_FMT VGPU_LOCK_DIR "/vgpu_%d.status"       /* B200 addition, written only while the former exists */
#define VGPU_READINGS_FMT VGPU_LOCK_DIR "/vgpu_%d.readings"   /* B200 addition: on-device utilisation readings (below) */
#define VGPU_TUNABLES_FILE VGPU_CFG_DIR "/b200.tunables"      /* B200 addition, optional: NAME=value lines from the control plane */
#define VGPU_SMUTIL_FILE VGPU_ROOT_DIR "/watcher/sm_util.config"
#define VGPU_SELF_FILE VGPU_ROOT_DIR "/driver/libvgpu-control.so"
#define VGPU_HOSTPROC_CGROUP_FMT VGPU_ROOT_DIR "/.host_proc/%d/cgroup"
#define VGPU_CLIENT_BIN VGPU_ROOT_DIR "/registry/device-client"
#define VGPU_LOCK_DIR "/tmp/.vgpu_lock"
#define VGPU_LOCK_FMT VGPU_LOCK_DIR "/vgpu_%d.lock"
#define VGPU_VMEM_DIR "/tmp/.vmem_node"
#define VGPU_VMEM_FILE VGPU_VMEM_DIR "/vmem_node.config"
#define VGPU_FAKE_UUID "GPU-00000000-0000-0000-0000-000000000000"

/* compatibility modes (hook.h:259-265); tested as (mode & K) == K */
enum {
  VGPU_MODE_HOST = 0,
  VGPU_MODE_CGROUPV1 = 1,
  VGPU_MODE_CGROUPV2 = 2,
  VGPU_MODE_OPEN_KERNEL = 100,
  VGPU_MODE_CLIENT = 200,
};

typedef struct {
  char uuid[VGPU_UUID_LEN];
  uint64_t total_memory; /* virtual cap seen by the tenant            */
  uint64_t real_memory;  /* physical share (== total unless oversold) */
  int32_t hard_core;
  int32_t soft_core;
  int32_t core_limit;
  int32_t hard_limit;
  int32_t memory_limit;
  int32_t memory_oversold;
  int32_t activate;
  int32_t _pad;
} vgpu_cfg_dev_t;

typedef struct {
  int32_t driver_major, driver_minor;
  char pod_uid[VGPU_UUID_LEN];
  char pod_name[VGPU_NAME_LEN];
  char pod_namespace[VGPU_NAME_LEN];
  char container_name[VGPU_NAME_LEN];
  vgpu_cfg_dev_t devices[VGPU_MAX_DEVICES]; /* index == HOST gpu index */
  int32_t compatibility_mode;
  int32_t sm_watcher;
  int32_t vmem_node;
  char reg_uuid[VGPU_UUID_LEN];
  int32_t _pad;
} vgpu_cfg_t;

/* NVML v1 process record as the reference reads it (nvml-subset.h:81-88) */
typedef struct {
  uint32_t pid;
  uint32_t _pad;
  uint64_t used_bytes;
} vgpu_proc_t;

/* nvmlProcessUtilizationSample_t (nvml-subset.h:498-505) */
typedef struct {
  uint32_t pid;
  uint32_t _pad;
  uint64_t ts_us;
  uint32_t sm, mem, enc, dec;
} vgpu_util_sample_t;

/* nvmlProcessInfoV2_t as embedded in sm_util.config (hook.h:200-205) */
typedef struct {
  uint32_t pid;
  uint32_t _pad;
  uint64_t used_bytes;
  uint32_t gi, ci;
} vgpu_proc_v2_t;

typedef struct {
  vgpu_util_sample_t samples[VGPU_MAX_PIDS];
  uint32_t samples_size;
  uint32_t _pad0;
  uint64_t last_seen_us;
  vgpu_proc_v2_t compute[VGPU_MAX_PIDS];
  uint32_t compute_size;
  uint32_t _pad1;
  vgpu_proc_v2_t graphics[VGPU_MAX_PIDS];
  uint32_t graphics_size;
  uint8_t lock_byte;
  uint8_t _pad2[3];
} vgpu_smutil_dev_t;

typedef struct {
  vgpu_smutil_dev_t devices[VGPU_MAX_DEVICES];
} vgpu_smutil_t;

typedef struct {
  int32_t pid;
  int32_t _pad;
  uint64_t used;
} vgpu_vmem_rec_t;

typedef struct {
  vgpu_vmem_rec_t processes[VGPU_MAX_PIDS];
  uint32_t processes_size;
  uint8_t lock_byte;
  uint8_t _pad[3];
} vgpu_vmem_dev_t;

typedef struct {
  vgpu_vmem_dev_t devices[VGPU_MAX_DEVICES];
} vgpu_vmem_t;

#define VGPU_STATIC_ASSERT(c, m) typedef char vgpu_sa_##m[(c) ? 1 : -1]
VGPU_STATIC_ASSERT(sizeof(vgpu_cfg_dev_t) == 96, cfg_dev);
VGPU_STATIC_ASSERT(sizeof(vgpu_cfg_t) == 1848, cfg);
VGPU_STATIC_ASSERT(offsetof(vgpu_cfg_t, devices) == 248, cfg_devs);
/* VGPU_LOCK_DIR/vgpu_<host index>.readings (B200 addition): a tenant whose limiter runs on an
 * on-device utilisation signal (stream queue-busy / per-SM probe) publishes the reading of every control period
 * here; the node's SM watcher (csrc/smwatcher.c --source device|mixed) turns the fresh ones into the per-process
 * samples of sm_util.config, so consumers (either interception library in external-watcher mode, the Go
 * device-monitor) get them without anybody polling nvmlDeviceGetProcessUtilization.  A tenant only knows its pid
 * inside its own pid namespace: it publishes that together with the namespace's inode and the watcher resolves
 * NVML's host pids through /proc/<pid>/ns/pid + the NSpid line of /proc/<pid>/status. */
#define VGPU_READINGS_SLOTS 1024
typedef struct {
  uint64_t owner;          /* (pid-namespace inode << 32) | pid inside that namespace; 0 = free slot (claimed by CAS) */
  uint64_t ts_us;          /* CLOCK_REALTIME of the publication, written last                                         */
  uint32_t sm_pct;         /* the controller's reading for this period (what NVML calls smUtil), 0..100               */
  uint32_t queue_busy_pct; /* its two ingredients, for diagnosis                                                      */
  uint32_t sm_active_pct;
  uint32_t seq;            /* publications by this owner                                                              */
} vgpu_reading_t;
typedef struct {
  vgpu_reading_t slots[VGPU_READINGS_SLOTS];
} vgpu_readings_t;
VGPU_STATIC_ASSERT(sizeof(vgpu_reading_t) == 32, reading);
VGPU_STATIC_ASSERT(sizeof(vgpu_readings_t) == 32768, readings);

VGPU_STATIC_ASSERT(offsetof(vgpu_cfg_t, compatibility_mode) == 1784, cfg_mode);
VGPU_STATIC_ASSERT(offsetof(vgpu_cfg_t, reg_uuid) == 1796, cfg_reg);
VGPU_STATIC_ASSERT(sizeof(vgpu_proc_t) == 16, proc);
VGPU_STATIC_ASSERT(sizeof(vgpu_util_sample_t) == 32, sample);
VGPU_STATIC_ASSERT(sizeof(vgpu_proc_v2_t) == 24, procv2);
VGPU_STATIC_ASSERT(sizeof(vgpu_smutil_dev_t) == 81952, smutil_dev);
VGPU_STATIC_ASSERT(offsetof(vgpu_smutil_dev_t, lock_byte) == 81948, smutil_lock);
VGPU_STATIC_ASSERT(sizeof(vgpu_smutil_t) == 1311232, smutil);
VGPU_STATIC_ASSERT(sizeof(vgpu_vmem_dev_t) == 16392, vmem_dev);
VGPU_STATIC_ASSERT(offsetof(vgpu_vmem_dev_t, lock_byte) == 16388, vmem_lock);
VGPU_STATIC_ASSERT(sizeof(vgpu_vmem_t) == 262272, vmem);

#ifdef __cplusplus
}
#endif
#endif
