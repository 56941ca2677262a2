/*
 * metrics.c - per-device counters (the reference's library/src/metrics.c:1-107 keeps lock-wait,
 * OOM, UVA-fallback, rate-limit and watcher counters and logs them at powers of two when
 * LOGGER_LEVEL >= 3; same idea, plus counters for the device kernels).
 */
#include "vgpu_internal.h"

static volatile uint64_t g_counters[VGPU_MAX_DEVICES][VM_COUNT];
static const char *g_names[VM_COUNT] = {"rate_gated",   "rate_fast",     "oom_total_limit", "oom_driver_return",
                                        "uva_fallback", "lock_timeout",  "quota_kernels",   "sampler_launches",
                                        "scrubbed_bytes", "watchdog_loans", "sampler_skipped",
                                        "spill_bytes", "spill_ns", "scrub_ns", "promote_bytes", "promote_ns",
                                        "slab_allocs", "slab_demotions"};

void vgpu_metric_add(int h, int which, uint64_t v) {
  if (h < 0 || h >= VGPU_MAX_DEVICES || which < 0 || which >= VM_COUNT) return;
  uint64_t total = __sync_add_and_fetch(&g_counters[h][which], v);
  if (which != VM_RATE_FAST && which != VM_SAMPLER_LAUNCHES && which != VM_SCRUBBED_BYTES && which != VM_SAMPLER_SKIPPED &&
      which < VM_SPILL_BYTES && vgpu_log_level() >= VL_INFO &&
      (total & (total - 1)) == 0)
    vgpu_log_emit(VL_INFO, __FILE__, __LINE__, "metric=%s host_device=%d total=%" PRIu64, g_names[which], h, total);
}

uint64_t vgpu_metric_get(int h, int which) {
  if (h < 0 || h >= VGPU_MAX_DEVICES || which < 0 || which >= VM_COUNT) return 0;
  return g_counters[h][which];
}
