/*
 * slabmode.c - VGPU_B200_SLAB=1: the spill path with real data movement.
 *
 * Reference behaviour (library/src/cuda_hook.c:93-116, :1344-1398): on an oversold device a
 * cuMemAlloc whose bytes no longer fit the physical share is rerouted to cuMemAllocManaged and its
 * size recorded in the UVA ledger; what lives in HBM afterwards is left to UVM's demand paging.
 *
 * Here, opt-in: cuMemAlloc / cuMemFree of an oversold, memory-limited device are served from
 * VMM-backed slabs (cuMemAddressReserve + cuMemCreate + cuMemMap), so the library - not a page-fault
 * handler - decides which slab is in HBM.  The ACCOUNTING is exactly the reference's: the quota
 * kernel's GPU / UVA / OOM decision, a ledger record for every allocation taken on the UVA path,
 * the same NVML-visible physical bytes.  The PLACEMENT differs: when the decision is "UVA" the
 * allocator demotes the coldest HBM-resident slab of the same size class instead - its pages are
 * staged into host memory with vgpu_spill_copy_kernel (TMA bulk copies) and remapped under the same
 * virtual address, the freed HBM is scrubbed with vgpu_clear_kernel (128-bit stores) and handed to
 * the new allocation.  New (hot) data lives in HBM, old data spills, and because victim and
 * newcomer have the same mapped size the device's physical usage and the ledger are
 * byte-identical to what the reference would report.  Frees restore the invariant
 *     #(UVA-accounted slabs in HBM) == #(GPU-accounted slabs on the host)      per size class
 * with at most one more copy (demote / promote of the partner slab).
 *
 * Bookkeeping lives in HBM (vgpu_vslab_kernel: free-slot scan, lookup, coldest-victim scan with
 * the placement flip in the same launch); the host half only holds what the driver needs, the
 * allocation handles.  Without a same-size victim the newcomer itself is host-backed (the
 * reference's placement).  Requests that are not whole multiples of the 2 MiB mapping granularity
 * stay on the plain path (their physical footprint would otherwise differ from the reference's).
 * Anything VMM cannot do here (no host-NUMA mappings, handles exhausted) falls back to the plain
 * path, loudly, once.
 */
#include "vgpu_internal.h"

#include <time.h>

typedef struct vgpu_slab_host {
  CUmemGenericAllocationHandle handle;
  CUdeviceptr va;
  uint64_t size;
} vgpu_slab_host; /* indexed by the device table's slot number */

static int g_mode = -1, g_broken;
static size_t g_gran;

int vgpu_slab_mode(void) {
  if (g_mode < 0) {
    const char *e = getenv("VGPU_B200_SLAB");
    g_mode = (e && (*e == '1' || *e == 't' || *e == 'T')) ? 1 : 0;
  }
  return g_mode && !g_broken;
}

static void give_up(const char *what, CUresult r) {
  if (!g_broken) VLOG(VL_ERROR, "slab mode disabled: %s failed (%d: %s); allocations take the plain path", what, r, vgpu_cu_err(r));
  g_broken = 1;
}

static int have_vmm(void) {
  return R.cuMemAddressReserve && R.cuMemAddressFree && R.cuMemCreate && R.cuMemRelease && R.cuMemMap && R.cuMemUnmap &&
         R.cuMemSetAccess && R.cuMemGetAllocationGranularity;
}

static void dev_prop(vcu_mem_alloc_prop_t *p, CUdevice dev) {
  memset(p, 0, sizeof *p);
  p->type = VCU_MEM_ALLOCATION_PINNED;
  p->location.type = VCU_MEM_LOCATION_DEVICE;
  p->location.id = dev;
}
static void host_prop(vcu_mem_alloc_prop_t *p, CUdevice dev) {
  int numa = -1;
  if (R.cuDeviceGetAttribute) R.cuDeviceGetAttribute(&numa, VCU_ATTR_HOST_NUMA_ID, dev);
  memset(p, 0, sizeof *p);
  p->type = VCU_MEM_ALLOCATION_PINNED;
  p->location.type = VCU_MEM_LOCATION_HOST_NUMA;
  p->location.id = numa >= 0 ? numa : 0;
}

static size_t granularity(CUdevice dev) {
  if (g_gran) return g_gran;
  vcu_mem_alloc_prop_t dp, hp;
  size_t gd = 0, gh = 0;
  dev_prop(&dp, dev);
  host_prop(&hp, dev);
  if (R.cuMemGetAllocationGranularity(&gd, &dp, VCU_MEM_GRANULARITY_MINIMUM) != CUDA_SUCCESS || !gd) gd = (size_t)2 << 20;
  if (R.cuMemGetAllocationGranularity(&gh, &hp, VCU_MEM_GRANULARITY_MINIMUM) != CUDA_SUCCESS || !gh) gh = gd;
  size_t g = gd > gh ? gd : gh;
  if (g % gd || g % gh) g = gd * gh; /* both are powers of two in practice */
  g_gran = g;
  return g;
}

/* ------------------------------------------------------------------ device table round trips */
static int vslab_call(vgpu_dev_rt *rt, vgpu_vslab_req_t *rq, vgpu_vslab_res_t *out) {
  uint32_t seq = ++rt->seq;
  if (!seq) seq = ++rt->seq;
  void *params[] = {&rt->vslab_d, rq, &rt->vs_res_d, &seq};
  if (vgpu_rt_launch(rt, rt->k_vslab, 1, 1024, 0, rt->q_stream, params) != CUDA_SUCCESS) return -1;
  /* the answer lands in pinned memory; a stream synchronise is fine here, nothing on this path is hot */
  if (R.cuStreamSynchronize(rt->q_stream) != CUDA_SUCCESS || rt->vs_res->seq_done != seq) return -1;
  *out = *rt->vs_res;
  return 0;
}

static int vslab_put(vgpu_dev_rt *rt, CUdeviceptr va, uint64_t bytes, uint64_t size, uint32_t state, uint32_t *slot) {
  vgpu_vslab_req_t rq = {VGPU_VSLAB_PUT, 0, 0, 0, 0, ++rt->vs_age, va, bytes, size, state, 0};
  vgpu_vslab_res_t res;
  if (vslab_call(rt, &rq, &res) || res.slot == 0xffffffffu) return -1;
  *slot = res.slot;
  return 0;
}

/* coldest slab of the size class whose placement bits are `want`; its bits become `set` */
static int vslab_pick(vgpu_dev_rt *rt, uint64_t size, uint32_t want, uint32_t set, vgpu_vslab_res_t *res) {
  vgpu_vslab_req_t rq = {VGPU_VSLAB_SCAN, VGPU_VS_UVA | VGPU_VS_DEV, want, VGPU_VS_DEV, set, 0, 0, 0, size, 0, 0};
  if (vslab_call(rt, &rq, res)) return -1;
  return res->slot == 0xffffffffu ? 1 : 0;
}

/* ------------------------------------------------------------------ VMM helpers */
static CUresult map_rw(CUdeviceptr va, size_t size, CUmemGenericAllocationHandle h, CUdevice dev) {
  CUresult r = R.cuMemMap(va, size, 0, h, 0);
  if (r != CUDA_SUCCESS) return r;
  vcu_mem_access_desc_t acc = {{VCU_MEM_LOCATION_DEVICE, dev}, VCU_MEM_ACCESS_READWRITE};
  r = R.cuMemSetAccess(va, size, &acc, 1);
  if (r != CUDA_SUCCESS) R.cuMemUnmap(va, size);
  return r;
}

static uint64_t timed_ns(vgpu_dev_rt *rt, struct timespec *t0) {
  /* device time between the two events when events are available, host clock otherwise */
  float ms = 0;
  if (rt->ev0 && rt->ev1 && R.cuEventSynchronize(rt->ev1) == CUDA_SUCCESS && R.cuEventElapsedTime(&ms, rt->ev0, rt->ev1) == CUDA_SUCCESS)
    return (uint64_t)((double)ms * 1e6);
  struct timespec t1;
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (uint64_t)(t1.tv_sec - t0->tv_sec) * 1000000000ull + (uint64_t)(t1.tv_nsec - t0->tv_nsec);
}

/* whole-device quiet point: nothing of the tenant may touch a slab while its backing is swapped */
static void tenant_quiesce(vgpu_dev_rt *rt) {
  vgpu_limiter_before_blocking_call(rt);
  vgpu_limiter_quiesce(rt);
  if (R.cuCtxSynchronize) R.cuCtxSynchronize();
  vgpu_limiter_resume(rt, 1);
}

/* Move the contents of slab `slot` into `to` (a handle of the other kind) and re-point its virtual
 * address there.  The old handle is returned (still allocated, unmapped).  kind: 0 demote (HBM ->
 * host, vgpu_spill_copy_kernel), 1 promote (host -> HBM). */
static CUresult swap_backing(vgpu_dev_rt *rt, CUdevice dev, uint32_t slot, CUmemGenericAllocationHandle to, int kind,
                             CUmemGenericAllocationHandle *old) {
  vgpu_slab_host *hs = &rt->vs_host[slot];
  CUdeviceptr stage = 0;
  CUresult r = R.cuMemAddressReserve(&stage, hs->size, g_gran, 0, 0);
  if (r != CUDA_SUCCESS) return r;
  r = map_rw(stage, hs->size, to, dev);
  if (r != CUDA_SUCCESS) { R.cuMemAddressFree(stage, hs->size); return r; }
  tenant_quiesce(rt);
  struct timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (rt->ev0) R.cuEventRecord(rt->ev0, rt->q_stream);
  r = vgpu_rt_spill(rt, stage, hs->va, hs->size, rt->q_stream);
  if (rt->ev1) R.cuEventRecord(rt->ev1, rt->q_stream);
  if (r == CUDA_SUCCESS) r = R.cuStreamSynchronize(rt->q_stream);
  if (r != CUDA_SUCCESS) {
    R.cuMemUnmap(stage, hs->size);
    R.cuMemAddressFree(stage, hs->size);
    return r;
  }
  uint64_t ns = timed_ns(rt, &t0);
  vgpu_metric_add(rt->host_index, kind ? VM_PROMOTE_BYTES : VM_SPILL_BYTES, hs->size);
  vgpu_metric_add(rt->host_index, kind ? VM_PROMOTE_NS : VM_SPILL_NS, ns);
  R.cuMemUnmap(hs->va, hs->size);
  R.cuMemUnmap(stage, hs->size);
  R.cuMemAddressFree(stage, hs->size);
  r = map_rw(hs->va, hs->size, to, dev);
  if (r != CUDA_SUCCESS) {
    /* the tenant's data is safe in `to` but its address is gone: nothing sane can follow */
    VLOG(VL_FATAL, "slab at 0x%llx could not be remapped after its contents were moved (%d: %s)", hs->va, r, vgpu_cu_err(r));
  }
  *old = hs->handle;
  hs->handle = to;
  return CUDA_SUCCESS;
}

static CUresult scrub_mapped(vgpu_dev_rt *rt, CUdeviceptr va, size_t size) {
  struct timespec t0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (rt->ev0) R.cuEventRecord(rt->ev0, rt->q_stream);
  CUresult r = vgpu_rt_clear(rt, va, size, rt->q_stream);
  if (rt->ev1) R.cuEventRecord(rt->ev1, rt->q_stream);
  if (r == CUDA_SUCCESS) r = R.cuStreamSynchronize(rt->q_stream);
  if (r == CUDA_SUCCESS) {
    vgpu_metric_add(rt->host_index, VM_SCRUB_NS, timed_ns(rt, &t0));
    vgpu_metric_add(rt->host_index, VM_SCRUBBED_BYTES, size);
  }
  return r;
}

/* ------------------------------------------------------------------ allocation */
static int ctx_in(vgpu_dev_rt *rt) {
  CUcontext cur = NULL;
  if (R.cuCtxGetCurrent && R.cuCtxGetCurrent(&cur) == CUDA_SUCCESS && cur == rt->ctx) return 0;
  return (R.cuCtxPushCurrent_v2 && R.cuCtxPushCurrent_v2(rt->ctx) == CUDA_SUCCESS) ? 1 : 0;
}
static void ctx_out(int pushed) {
  CUcontext dummy;
  if (pushed) R.cuCtxPopCurrent_v2(&dummy);
}

CUresult vgpu_slab_alloc(vgpu_dev_rt *rt, CUdevice dev, int path, CUdeviceptr *dptr, size_t bytes, int *recorded_uva) {
  *recorded_uva = 0;
  if (!vgpu_slab_mode() || !rt || !bytes) return CUDA_ERROR_NOT_SUPPORTED;
  if (!have_vmm()) { give_up("resolving the VMM entry points", CUDA_ERROR_NOT_FOUND); return CUDA_ERROR_NOT_SUPPORTED; }
  pthread_mutex_lock(&rt->q_mu);
  int pushed = ctx_in(rt);
  CUresult r = CUDA_ERROR_NOT_SUPPORTED;
  if (!rt->vs_host) rt->vs_host = (vgpu_slab_host *)calloc(VGPU_VSLAB_SLOTS, sizeof(vgpu_slab_host));
  if (!rt->vs_host) goto out;
  if (!rt->ev0 && R.cuEventCreate && R.cuEventRecord && R.cuEventSynchronize && R.cuEventElapsedTime) {
    if (R.cuEventCreate(&rt->ev0, 0) != CUDA_SUCCESS) rt->ev0 = NULL;
    if (!rt->ev0 || R.cuEventCreate(&rt->ev1, 0) != CUDA_SUCCESS) rt->ev0 = rt->ev1 = NULL;
  }
  const size_t gran = granularity(dev);
  /* Only requests that are whole multiples of the mapping granularity (2 MiB) become slabs: a VMM mapping cannot be
   * smaller, so anything else would occupy more physical memory than the reference's cuMemAlloc of the same bytes and the
   * NVML-visible `used` - hence every later GPU / UVA / OOM decision - would drift from the reference's (found by the
   * offline fuzz sweep: 1-byte .. 3 MiB+17 requests).  Those take the plain path: the reference's placement, exactly. */
  if (bytes % gran) { r = CUDA_ERROR_NOT_SUPPORTED; goto out; }
  const size_t size = bytes;
  vcu_mem_alloc_prop_t dp, hp;
  dev_prop(&dp, dev);
  host_prop(&hp, dev);

  CUdeviceptr va = 0;
  r = R.cuMemAddressReserve(&va, size, gran, 0, 0);
  if (r != CUDA_SUCCESS) { give_up("cuMemAddressReserve", r); r = CUDA_ERROR_NOT_SUPPORTED; goto out; }

  CUmemGenericAllocationHandle h = 0;
  uint32_t state = 0;
  if (path == VGPU_PATH_GPU) {
    r = R.cuMemCreate(&h, size, &dp, 0);
    if (r == CUDA_ERROR_OUT_OF_MEMORY) path = VGPU_PATH_UVA; /* the driver is out of HBM: same retry as the reference (:1372-1386) */
    else if (r != CUDA_SUCCESS) { R.cuMemAddressFree(va, size); give_up("cuMemCreate(device)", r); r = CUDA_ERROR_NOT_SUPPORTED; goto out; }
    else state = VGPU_VS_DEV;
  }
  if (path == VGPU_PATH_UVA) {
    /* the spill decision: coldest HBM-resident, GPU-accounted slab of this size class */
    vgpu_vslab_res_t victim;
    int none = vslab_pick(rt, size, VGPU_VS_DEV, 0, &victim);
    CUmemGenericAllocationHandle hostbuf = 0;
    r = R.cuMemCreate(&hostbuf, size, &hp, 0);
    if (r != CUDA_SUCCESS) {
      if (none == 0) { /* undo the flip */
        vgpu_vslab_req_t undo = {VGPU_VSLAB_SCAN, 0, 0, VGPU_VS_DEV, VGPU_VS_DEV, 0, victim.dptr, 0, 0, 0, 0};
        vgpu_vslab_res_t dropped;
        vslab_call(rt, &undo, &dropped);
      }
      R.cuMemAddressFree(va, size);
      give_up("cuMemCreate(host NUMA)", r);
      r = CUDA_ERROR_NOT_SUPPORTED;
      goto out;
    }
    if (none == 0) {
      CUmemGenericAllocationHandle freed = 0;
      r = swap_backing(rt, dev, victim.slot, hostbuf, 0, &freed);
      if (r != CUDA_SUCCESS) {
        vgpu_vslab_req_t undo = {VGPU_VSLAB_SCAN, 0, 0, VGPU_VS_DEV, VGPU_VS_DEV, 0, victim.dptr, 0, 0, 0, 0};
        vgpu_vslab_res_t dropped;
        vslab_call(rt, &undo, &dropped);
        R.cuMemRelease(hostbuf);
        R.cuMemAddressFree(va, size);
        give_up("demoting a slab", r);
        r = CUDA_ERROR_NOT_SUPPORTED;
        goto out;
      }
      vgpu_metric_add(rt->host_index, VM_SLAB_DEMOTIONS, 1);
      h = freed; /* the victim's HBM now backs the newcomer - after it has been scrubbed */
      state = VGPU_VS_UVA | VGPU_VS_DEV;
    } else {
      h = hostbuf; /* no victim of this size: the newcomer itself lives on the host (the reference's placement) */
      state = VGPU_VS_UVA;
    }
    *recorded_uva = 1;
  }
  r = map_rw(va, size, h, dev);
  if (r == CUDA_SUCCESS && (state & VGPU_VS_DEV) && (state & VGPU_VS_UVA)) r = scrub_mapped(rt, va, size);
  uint32_t slot = 0;
  if (r == CUDA_SUCCESS && vslab_put(rt, va, bytes, size, state, &slot)) r = CUDA_ERROR_OUT_OF_MEMORY; /* table full */
  if (r != CUDA_SUCCESS) {
    R.cuMemUnmap(va, size);
    R.cuMemRelease(h);
    R.cuMemAddressFree(va, size);
    *recorded_uva = 0;
    goto out;
  }
  rt->vs_host[slot].handle = h;
  rt->vs_host[slot].va = va;
  rt->vs_host[slot].size = size;
  *dptr = va;
  vgpu_metric_add(rt->host_index, VM_SLAB_ALLOCS, 1);
out:
  ctx_out(pushed);
  pthread_mutex_unlock(&rt->q_mu);
  return r;
}

/* ------------------------------------------------------------------ free */
int vgpu_slab_free(vgpu_dev_rt *rt, CUdeviceptr dptr, CUresult *out, int *was_uva, uint64_t *bytes) {
  if (!rt || !rt->vs_host || !dptr) return 0;
  pthread_mutex_lock(&rt->q_mu);
  int pushed = ctx_in(rt);
  int handled = 0;
  vgpu_vslab_req_t rq = {VGPU_VSLAB_TAKE, 0, 0, 0, 0, 0, dptr, 0, 0, 0, 0};
  vgpu_vslab_res_t rec;
  if (vslab_call(rt, &rq, &rec) == 0 && rec.slot != 0xffffffffu) {
    handled = 1;
    vgpu_slab_host hs = rt->vs_host[rec.slot];
    memset(&rt->vs_host[rec.slot], 0, sizeof hs);
    *was_uva = (rec.state & VGPU_VS_UVA) != 0;
    *bytes = rec.bytes;
    CUdevice dev = rt->cuda_dev;
    tenant_quiesce(rt); /* cuMemFree semantics: pending work that may still use the buffer finishes first */
    R.cuMemUnmap(hs.va, hs.size);
    R.cuMemAddressFree(hs.va, hs.size);
    const int uva = (rec.state & VGPU_VS_UVA) != 0, in_hbm = (rec.state & VGPU_VS_DEV) != 0;
    CUmemGenericAllocationHandle spare = hs.handle;
    if (uva && in_hbm) {
      /* its partner - a GPU-accounted slab that was demoted - comes home into the HBM this one held */
      vgpu_vslab_res_t p;
      if (vslab_pick(rt, hs.size, 0, VGPU_VS_DEV, &p) == 0) {
        CUmemGenericAllocationHandle hostbuf = 0;
        if (swap_backing(rt, dev, p.slot, hs.handle, 1, &hostbuf) == CUDA_SUCCESS) spare = hostbuf;
        else {
          vgpu_vslab_req_t undo = {VGPU_VSLAB_SCAN, 0, 0, VGPU_VS_DEV, 0, 0, p.dptr, 0, 0, 0, 0};
          vgpu_vslab_res_t dropped;
          vslab_call(rt, &undo, &dropped);
        }
      }
    } else if (!uva && !in_hbm) {
      /* a demoted GPU-accounted slab goes away: the UVA-accounted slab that took its HBM gives it up */
      vgpu_vslab_res_t p;
      if (vslab_pick(rt, hs.size, VGPU_VS_UVA | VGPU_VS_DEV, 0, &p) == 0) {
        CUmemGenericAllocationHandle freed = 0;
        if (swap_backing(rt, dev, p.slot, hs.handle, 0, &freed) == CUDA_SUCCESS) {
          spare = freed;
          vgpu_metric_add(rt->host_index, VM_SLAB_DEMOTIONS, 1);
        } else {
          vgpu_vslab_req_t undo = {VGPU_VSLAB_SCAN, 0, 0, VGPU_VS_DEV, VGPU_VS_DEV, 0, p.dptr, 0, 0, 0, 0};
          vgpu_vslab_res_t dropped;
          vslab_call(rt, &undo, &dropped);
        }
      }
    }
    *out = R.cuMemRelease(spare);
  }
  ctx_out(pushed);
  pthread_mutex_unlock(&rt->q_mu);
  return handled;
}

void vgpu_slab_forget(vgpu_dev_rt *rt) {
  if (!rt) return;
  free(rt->vs_host);
  rt->vs_host = NULL;
  rt->ev0 = rt->ev1 = NULL;
}
