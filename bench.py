#!/usr/bin/env python
"""bench.py - BASELINE.json's metric for the cuLaunchKernel gate + memory-cap hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[1] - one tenant per B200, 25 % cores /
4 GiB cap, an empty-kernel <<<1,1>>> cuLaunchKernel storm issued by a C tenant
(tests/harness/storm.c) that runs with libvgpu-control.so LD_PRELOADed, i.e. through the
reference-facing symbol surface.  A "step" is one batch of PER_STEP launches followed by a
device synchronise.

  value     intercepted launches / second over the K timed steps, all ranks (N tenants, one per
            GPU, weak scaling), timed on the device with CUDA events around each step, max over
            ranks (the limiter state - token bucket, sampler - is resident in HBM throughout).
  e2e       the same K steps on the tenant's HOST clock (launch calls through the LD_PRELOADed
            hook + the per-step device synchronise); host<->device bytes are what the hook itself
            moves over PCIe per launch (ticket + launch sequence in pinned memory read by the
            sampler, completion markers written back).  `tenant_process_life_s` additionally
            reports the whole process life (exec, dlopen, bring-up, storm, teardown).
  roofline  the spill-copy kernel (TMA bulk HBM->HBM staging of spilled pages, the dominant
            device kernel of the memory path): algorithmic bytes 2 x 1 GiB per launch over the
            CUDA-event time on the launching stream, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference
            the UNMODIFIED reference library (oracle/_ref/libvgpu-control.so, built from
            /root/reference/library by oracle/Makefile) preloaded into the same tenant on the
            same box: its limiter is the CPU watcher thread (NVML poll + nanosleep gate).

With torchrun (N > 1) every rank drives its own GPU's tenant; the only collective is the
cross-tenant rebalance vector (NCCL all_gather of {quota, achieved, slack} per GPU), timed
separately and reported as `rebalance_us`.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PER_STEP = 200_000
CORE_LIMIT = 25
MEM_LIMIT = "4g"
FAKE = "GPU-00000000-0000-0000-0000-000000000000"


def build_everything():
    import helpers
    helpers.build_all()
    return helpers


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def gpu_uuids():
    out = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader"], capture_output=True, text=True)
    return [l.strip() for l in out.stdout.splitlines() if l.strip()]


def tenant_env(H, lib, gpu, uuids, sandbox):
    """Environment of one tenant container: what the device plugin's Allocate() injects
    (reference pkg/deviceplugin/vgpu/vnum_plugin.go:568-757), here in its env form."""
    env = {k: v for k, v in os.environ.items() if not k.startswith(("CUDA_", "MANAGER_", "VGPU_", "LD_PRELOAD"))}
    vis = [FAKE] * 16
    vis[gpu] = uuids[gpu]
    env.update({
        "CUDA_VISIBLE_DEVICES": str(gpu),
        "MANAGER_COMPATIBILITY_MODE": "0",
        "MANAGER_VISIBLE_DEVICES": ",".join(vis),
        "CUDA_CORE_LIMIT_%d" % gpu: str(CORE_LIMIT),
        "CUDA_MEM_LIMIT_%d" % gpu: MEM_LIMIT,
        "LOGGER_LEVEL": "1",
        "VGPU_REDIRECT": ":".join(["/etc/vgpu-manager=%s/etc/vgpu-manager" % sandbox, "/tmp/.vgpu_lock=%s/lock" % sandbox,
                                   "/tmp/.vmem_node=%s/vmem" % sandbox]),
        "LD_PRELOAD": H.REDIRECT + " " + lib,
    })
    return env


def run_tenant(H, lib, gpu, uuids, steps, warmup, per_step, max_seconds=0.0):
    sandbox = tempfile.mkdtemp(prefix="vgpu_bench_")
    for d in ("etc/vgpu-manager/config", "lock", "vmem"):
        os.makedirs(os.path.join(sandbox, d), exist_ok=True)
    cmd = [H.STORM, "--steps", str(steps), "--warmup", str(warmup), "--per-step", str(per_step), "--device", "0",
           "--host-index", str(gpu)]
    if max_seconds:
        cmd += ["--max-seconds", str(max_seconds)]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, env=tenant_env(H, lib, gpu, uuids, sandbox), capture_output=True, text=True,
                       timeout=max(240.0, 4 * max_seconds))
    life = time.perf_counter() - t0
    shutil.rmtree(sandbox, ignore_errors=True)
    if r.returncode != 0 or not r.stdout.strip():
        raise RuntimeError("tenant failed rc=%d\n%s\n%s" % (r.returncode, r.stdout[-2000:], r.stderr[-3000:]))
    d = json.loads(r.stdout.strip().splitlines()[-1])
    d["life_s"] = life
    d["stderr_tail"] = r.stderr[-400:]
    return d


def spill_roofline(lib_path, peaks):
    """Time vgpu_spill_copy_kernel in-process: 1 GiB HBM->HBM (>> 126 MB L2), CUDA events on the
    launching stream, 3 warm-ups, 10 timed launches."""
    import torch
    from vgpu_manager_b200 import B200Library
    torch.zeros(1, device="cuda")
    uuid = "GPU-" + str(torch.cuda.get_device_properties(0).uuid)
    lib = B200Library(path=lib_path, env={"MANAGER_VISIBLE_DEVICES": uuid, "MANAGER_COMPATIBILITY_MODE": "0"})
    lib.attach()
    n = 1 << 30
    src = torch.randint(0, 255, (n,), dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    stream = torch.cuda.current_stream()
    for _ in range(3):
        lib.spill_copy(dst.data_ptr(), src.data_ptr(), n, stream.cuda_stream)
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        lib.spill_copy(dst.data_ptr(), src.data_ptr(), n, stream.cuda_stream)
        e1.record(stream)
        e1.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    assert torch.equal(src, dst)
    avg = sum(times) / len(times)
    # the clear kernel, same method (N bytes written)
    ctimes = []
    for i in range(13):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        lib.clear(dst.data_ptr(), n, stream.cuda_stream)
        e1.record(stream)
        e1.synchronize()
        if i >= 3:
            ctimes.append(e0.elapsed_time(e1) * 1e-3)
    cavg = sum(ctimes) / len(ctimes)
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = 2 * n / avg / 1e9
    return {"bound": "hbm", "kernel": "vgpu_spill_copy_kernel", "achieved": round(achieved, 1), "peak": peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)",
            "unit": "GB/s", "frac": round(achieved / peak, 4),
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this size
            # (ncu --set full, round 1: profiles/ncu_spill_copy_r1.txt: 1.073823 + 1.026727 GB)
            "traffic": 2100550000,
            "algorithmic_bytes_per_launch": 2 * n, "avg_launch_ms": round(avg * 1e3, 4), "best_launch_ms": round(min(times) * 1e3, 4),
            "launches_timed": len(times),
            "clear": {"kernel": "vgpu_clear_kernel", "achieved": round(n / cavg / 1e9, 1), "unit": "GB/s",
                      "frac": round(n / cavg / 1e9 / peak, 4), "algorithmic_bytes_per_launch": n,
                      "avg_launch_ms": round(cavg * 1e3, 4)}}, 10 + 3 + 13


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-step", type=int, default=PER_STEP)
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--roofline-only", action="store_true",
                    help="only the in-process bandwidth kernels (the form profiled under ncu: a tenant that runs with the "
                         "interposer preloaded cannot also run under ncu's - its gate waits deadlock with kernel serialisation)")
    args = ap.parse_args()
    if args.roofline_only:
        import helpers
        helpers.build_all()
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        roof, n = spill_roofline(helpers.NEW_SO, peaks)
        print(json.dumps({"roofline": roof, "gpu_launches": n}))
        return
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    if args.impl == "reference" and distributed:
        # the reference arm is a CPU path: rank 0 alone times it, the other ranks leave at once
        if rank != 0:
            return
        distributed = False

    import torch
    import torch.distributed as dist
    if distributed:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # one rank builds (make / nvcc are not re-entrant on the same tree), the others wait
    if rank == 0:
        H = build_everything()
    if distributed:
        dist.barrier()
    if rank != 0:
        import helpers as H
    lib = H.NEW_SO if args.impl == "b200" else H.REF_SO
    if args.impl == "reference" and not os.path.exists(H.REF_SO):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvgpu-control.so was not built"}))
        return
    if args.impl == "b200" and not os.path.exists(H.NEW_SO):
        raise SystemExit("libvgpu-control.so (with its sm_100a image) is missing - no CPU fallback exists")
    uuids = gpu_uuids()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # the reference arm throttles on the CPU in 10 ms sleeps: bound each of its steps
    max_seconds = 0.0
    per_step = args.per_step
    if args.impl == "reference":
        max_seconds = 20.0 * args.steps / 5.0

    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    t_region0 = time.perf_counter()
    res = run_tenant(H, lib, local_rank, uuids, args.steps, args.warmup, per_step, max_seconds)
    barrier()
    region_s = time.perf_counter() - t_region0
    clk = clocks.stop()

    # timed K steps on the device (events inside the tenant); max over ranks
    dev_s = res["device_s"] if res.get("device_s", 0) > 0 else res["wall_s"]
    vals = torch.tensor([dev_s, float(res["launches"]), res["wall_s"], float(res["p50_ns"]), float(res["p99_ns"]),
                         float(res.get("sampler_launches", 0)), float(res.get("gated_launches", 0))],
                        dtype=torch.float64, device="cuda")
    rebalance_us = None
    if distributed:
        gathered = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(gathered, vals)
        # cross-tenant rebalance vector: {gpu, quota, achieved launches/s, gated fraction}
        from vgpu_manager_b200.multi import TenantReport, all_gather_reports, rebalance
        rep = TenantReport(local_rank, CORE_LIMIT, res["launches"] / dev_s,
                           res.get("gated_launches", 0) / max(res["launches"], 1))
        table = all_gather_reports(dist, torch, rep, torch.device("cuda", local_rank))
        plan = rebalance(table)
        # time the collective itself (pre-allocated buffers, CUDA events, max over ranks)
        vec = torch.tensor(rep.as_vector(), dtype=torch.float32, device="cuda")
        bufs = [torch.zeros_like(vec) for _ in range(world)]
        for _ in range(5):
            dist.all_gather(bufs, vec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_gather(bufs, vec)
        e1.record()
        e1.synchronize()
        rb = torch.tensor([e0.elapsed_time(e1) * 1e3 / 20], dtype=torch.float64, device="cuda")
        dist.all_reduce(rb, op=dist.ReduceOp.MAX)
        rebalance_us = float(rb.item())
        rows = [g.tolist() for g in gathered]
    else:
        rows = [vals.tolist()]

    roof, own_launches = None, 0
    cpu_base = None
    if rank == 0 and args.impl == "b200" and not args.no_roofline:
        roof, own_launches = spill_roofline(H.NEW_SO, peaks)
        if os.path.exists(H.REF_SO):
            # bounded sample of the same workload under the reference's CPU watcher path
            ref = run_tenant(H, H.REF_SO, local_rank, uuids, 1, 1, per_step, max_seconds=20.0)
            rdev = ref["device_s"] if ref.get("device_s", 0) > 0 else ref["wall_s"]
            cpu_base = {"value": round(ref["launches"] / rdev, 1), "unit": "launches/s", "cores": 1, "kind": "reference",
                        "host_cores_on_box": os.cpu_count(),
                        "sample": "1 warm-up + 1 timed step of up to %d launches (20 s cap, %d done) under "
                                  "oracle/_ref/libvgpu-control.so, same 25%%/4GiB tenant env" % (per_step, ref["launches"]),
                        "p50_hook_ns": ref["p50_ns"], "p99_hook_ns": ref["p99_ns"], "max_hook_ns": ref["max_ns"]}
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    t_max = max(r[0] for r in rows)
    total = sum(r[1] for r in rows)
    life_max = max(r[2] for r in rows)
    value = total / t_max
    line = {
        "metric": "intercepted cuLaunchKernel/sec under 25% core / 4 GiB cap (p50 hook latency alongside)",
        "value": round(value, 1), "unit": "launches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(t_max / max(res["steps"], 1) * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "configs[1]: 1 tenant per B200, 25%% cores / 4 GiB cap, empty-kernel <<<1,1>>> "
                               "cuLaunchKernel storm, %d launches per step, sync per step" % per_step,
                   "per_step_launches": per_step, "core_limit_pct": CORE_LIMIT, "mem_limit": MEM_LIMIT,
                   "l2_policy": "storm has no data reuse; roofline copy uses 1 GiB buffers (> 126 MB L2)",
                   "tenants": world if args.impl == "b200" else 1, "impl_library": os.path.relpath(lib, ROOT)},
        "p50_hook_ns": max(r[3] for r in rows), "p99_hook_ns": max(r[4] for r in rows),
        "clocks": clk,
        "e2e": {"value": round(total / life_max, 1), "unit": "launches/s",
                "h2d_bytes_per_step": 16 * per_step if args.impl == "b200" else 0,
                "d2h_bytes_per_step": 8 * per_step // 256 if args.impl == "b200" else 0,
                "what": "same K steps on the tenant's HOST clock around launch calls + device sync, through the "
                        "LD_PRELOADed hook; h2d = ticket + sequence words the hook publishes per launch in pinned "
                        "memory (read by the sampler over PCIe), d2h = completion markers (one per 256 launches)"},
        "tenant_process_life_s": round(res["life_s"], 3),
        "gpu_launches": int(sum(r[5] for r in rows)) + own_launches,
        "gated_launches": int(sum(r[6] for r in rows)),
        "limiter_rank0": res.get("limiter"),
        "truncated": bool(res.get("truncated", 0)),
        "region_wall_s": round(region_s, 3),
    }
    if rebalance_us is not None:
        line["rebalance_us"] = round(rebalance_us, 2)
        line["rebalance_plan_pct"] = {str(k): round(v, 1) for k, v in plan.items()}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["gpu_launches"] = 0
        line["cpu_baseline"] = {"value": line["value"], "unit": "launches/s", "cores": 1, "kind": "reference",
                                "host_cores_on_box": os.cpu_count(),
                                "sample": "%d timed steps of up to %d launches (%.0f s cap) under oracle/_ref" %
                                          (res["steps"], per_step, max_seconds)}
        line["e2e"] = {"value": line["value"], "unit": "launches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    else:
        line["roofline"] = roof
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))


if __name__ == "__main__":
    main()
