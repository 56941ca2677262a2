#!/usr/bin/env python
"""bench.py - BASELINE.json's metric for the cuLaunchKernel gate + memory-cap hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload):
  N = 1   BASELINE.json configs[1] - one tenant on one B200, 25 % cores / 4 GiB cap,
  N > 1   BASELINE.json configs[4] - one tenant per B200, 50 % cores each (+ the same 4 GiB cap),
an empty-kernel <<<1,1>>> cuLaunchKernel storm issued by a C tenant (tests/harness/storm.c) that
runs in its own mount namespace (own /etc/vgpu-manager, /tmp/.vgpu_lock, /tmp/.vmem_node, like a
container) with libvgpu-control.so LD_PRELOADed, i.e. through the reference-facing symbol surface.
A "step" is one batch of PER_STEP launches followed by a device synchronise.  Both arms (--impl
b200 / reference) run the same tenants, one per rank/GPU, and are aggregated the same way.

  value     intercepted launches / second over the K timed steps, all ranks (N tenants, one per
            GPU, weak scaling), timed on the device with CUDA events around each step, max over
            ranks (the limiter state - token bucket, controller - is resident in HBM throughout).
  e2e       the same K steps on the tenant's HOST clock (launch calls through the LD_PRELOADed
            hook + the per-step device synchronise); host<->device bytes are what the hook itself
            moves over PCIe per step.
  added_p50_ns / added_p99_ns
            hook latency minus the bare driver's (same tenant, no preload, same run, rank 0).
  achieved_util_pct
            mean `utilization.gpu` (nvidia-smi, 200 ms) over the timed region - printed for both
            arms so that a higher launch rate cannot be bought with a looser throttle.
  alloc_path
            config-4 allocator storm (8 GiB cap oversold 4x, ledger on): cuMemAlloc/cuMemFree p50/p99
            through the hooks, both libraries, same run (rank 0, N = 1).
  roofline  bandwidth kernels of the memory path: CUDA-event time on the launching
            stream against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference
            the UNMODIFIED reference library (oracle/_ref/libvgpu-control.so, built from
            /root/reference/library by oracle/Makefile) preloaded into the same tenants on the
            same box: its limiter is the CPU watcher thread (NVML poll + nanosleep gate).

With torchrun (N > 1) every rank drives its own GPU's tenant; the only collective is the
cross-tenant rebalance (all_gather_into_tensor of {quota, utilisation, gated} per GPU -> plan ->
vgpu_b200_set_limits), see vgpu_manager_b200/multi.py.
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
MEM_LIMIT = "4g"
FAKE = "GPU-00000000-0000-0000-0000-000000000000"


def build_everything():
    import helpers
    helpers.build_all()
    return helpers


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons / utilisation during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "utilization.gpu")

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

        def num(r, i):
            try:
                return float(r[i])
            except Exception:
                return None
        sm = sorted(int(v) for v in (num(r, 1) for r in self.rows if len(r) > 2) if v is not None)
        mx = [int(v) for v in (num(r, 2) for r in self.rows if len(r) > 2) if v is not None]
        util = [v for v in (num(r, 9) for r in self.rows if len(r) > 9) if v is not None]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(nm)
        return ({"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                 "reasons": sorted(reasons), "samples": len(self.rows)},
                round(sum(util) / len(util), 1) if util else None)


def gpu_uuids():
    out = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader"], capture_output=True, text=True)
    return [l.strip() for l in out.stdout.splitlines() if l.strip()]


_MOUNTNS = None


def have_mountns():
    """Tenants get private contract directories the way containers do: a mount namespace with bind
    mounts.  Where that is not permitted the path-redirect shim of the test tree is the fallback
    (reported in config.isolation)."""
    global _MOUNTNS
    if _MOUNTNS is None:
        d = tempfile.mkdtemp(prefix="vgpu_ns_probe_")
        os.makedirs(os.path.join(d, "a"))
        os.makedirs(os.path.join(d, "b"))
        r = subprocess.run(["unshare", "-m", "sh", "-c", "mount --bind %s/a %s/b" % (d, d)], capture_output=True)
        _MOUNTNS = r.returncode == 0
        shutil.rmtree(d, ignore_errors=True)
    return _MOUNTNS


def tenant_env(H, lib, gpu, uuids, sandbox, core_limit, extra=None):
    """Environment of one tenant container: what the device plugin's Allocate() injects
    (reference pkg/deviceplugin/vgpu/vnum_plugin.go:568-757), here in its env form."""
    env = {k: v for k, v in os.environ.items() if not k.startswith(("CUDA_", "MANAGER_", "VGPU_", "LD_PRELOAD"))}
    vis = [FAKE] * 16
    vis[gpu] = uuids[gpu]
    env.update({
        "CUDA_VISIBLE_DEVICES": str(gpu),
        "MANAGER_COMPATIBILITY_MODE": "0",
        "MANAGER_VISIBLE_DEVICES": ",".join(vis),
        "CUDA_MEM_LIMIT_%d" % gpu: MEM_LIMIT,
        "LOGGER_LEVEL": "1",
    })
    if core_limit:
        env["CUDA_CORE_LIMIT_%d" % gpu] = str(core_limit)
    if extra:
        env.update({k.replace("%d", str(gpu)): v for k, v in extra.items()})
    pre = []
    if not have_mountns():
        env["VGPU_REDIRECT"] = ":".join(["/etc/vgpu-manager=%s/etc/vgpu-manager" % sandbox, "/tmp/.vgpu_lock=%s/lock" % sandbox,
                                         "/tmp/.vmem_node=%s/vmem" % sandbox])
        pre.append(H.REDIRECT)
    if lib:
        pre.append(lib)
    if pre:
        env["LD_PRELOAD"] = " ".join(pre)
    return env


def in_container(cmd, sandbox):
    if not have_mountns():
        return cmd
    script = ("mkdir -p /etc/vgpu-manager /tmp/.vgpu_lock /tmp/.vmem_node && "
              "mount --bind {sb}/etc/vgpu-manager /etc/vgpu-manager && mount --bind {sb}/lock /tmp/.vgpu_lock && "
              "mount --bind {sb}/vmem /tmp/.vmem_node && exec \"$@\"").format(sb=sandbox)
    # LD_PRELOAD must only reach the tenant, not unshare/sh/mount: it is re-exported right before exec
    return ["unshare", "-m", "sh", "-c", "P=\"$VGPU_TENANT_PRELOAD\"; unset VGPU_TENANT_PRELOAD; " +
            script.replace('exec "$@"', 'LD_PRELOAD="$P" exec "$@"'), "tenant"] + cmd


def run_in_tenant(H, cmd, lib, gpu, uuids, core_limit, extra=None, timeout=300.0, agent=None):
    """`agent(sandbox)` may return a started thread (the GPU's node agent) that lives beside the tenant."""
    sandbox = tempfile.mkdtemp(prefix="vgpu_bench_")
    for d in ("etc/vgpu-manager/config", "lock", "vmem"):
        os.makedirs(os.path.join(sandbox, d), exist_ok=True)
    env = tenant_env(H, lib, gpu, uuids, sandbox, core_limit, extra)
    if have_mountns():
        env["VGPU_TENANT_PRELOAD"] = env.pop("LD_PRELOAD", "")
    side = agent(sandbox) if agent else None
    t0 = time.perf_counter()
    r = subprocess.run(in_container(cmd, sandbox), env=env, capture_output=True, text=True, timeout=timeout,
                       preexec_fn=H.pin_to(H.gpu_local_cpus(gpu)))
    life = time.perf_counter() - t0
    if side is not None:
        side.join(timeout=120)
    shutil.rmtree(sandbox, ignore_errors=True)
    if r.returncode != 0 or not r.stdout.strip():
        raise RuntimeError("tenant failed rc=%d\n%s\n%s" % (r.returncode, r.stdout[-2000:], r.stderr[-3000:]))
    d = json.loads(r.stdout.strip().splitlines()[-1])
    d["life_s"] = life
    d["stderr_tail"] = r.stderr[-1500:]
    return d


def run_tenant(H, lib, gpu, uuids, steps, warmup, per_step, core_limit, max_seconds=0.0, agent=None):
    cmd = [H.STORM, "--steps", str(steps), "--warmup", str(warmup), "--per-step", str(per_step), "--device", "0",
           "--host-index", str(gpu)]
    if max_seconds:
        cmd += ["--max-seconds", str(max_seconds)]
    return run_in_tenant(H, cmd, lib, gpu, uuids, core_limit, timeout=max(240.0, 4 * max_seconds), agent=agent)


def run_allocstorm(H, lib, gpu, uuids, n=1500, nbytes=1 << 20):
    """Config-4 cap (8 GiB oversold 4x, ledger on), allocator storm through the hooks."""
    extra = {"CUDA_MEM_LIMIT_%d": "8g", "CUDA_MEM_RATIO_%d": "4", "VMEMORY_NODE_ENABLED": "true"}
    cmd = [os.path.join(H.BUILD, "allocstorm"), "--n", str(n), "--bytes", str(nbytes), "--device", "0"]
    d = run_in_tenant(H, cmd, lib, gpu, uuids, 0, extra)
    return {k: d[k] for k in ("pairs", "pairs_per_s", "alloc_p50_ns", "alloc_p99_ns", "free_p50_ns", "free_p99_ns", "fails")}


def slab_leg(H, gpu, uuids, peaks):
    """The memory path with real data movement, measured from INSIDE the hooks: config 4's shape scaled x4 (32 GiB cap
    over 8 GiB physical, so that several 1 GiB slabs spill), VGPU_B200_SLAB=1, sixteen 1 GiB cuMemAlloc calls.  Every "UVA" decision of the quota kernel makes the
    cuMemAlloc hook demote the coldest HBM slab to host memory (vgpu_spill_copy_kernel, TMA bulk copies over PCIe),
    scrub the freed HBM (vgpu_clear_kernel, 128-bit stores - the HBM-bound kernel of this path) and hand it to the
    new allocation.  Times are CUDA events recorded by the library around the kernels it launched."""
    gib = 1 << 30
    script = "init 0\nnvmlinit 0\n" + "".join("alloc %d\nfill %d %d %d\n" % (gib, i, gib, 17 + i) for i in range(16))
    script += "".join("check %d %d %d\n" % (i, gib, 17 + i) for i in range(16)) + "nvmlinfo\nledger 0\nslabstats 0\n"
    extra = {"CUDA_MEM_LIMIT_%d": "32g", "CUDA_MEM_RATIO_%d": "4", "VMEMORY_NODE_ENABLED": "true", "VGPU_B200_SLAB": "1",
             "SCENARIO_LEDGER": "/tmp/.vmem_node/vmem_node.config"}
    sandbox = tempfile.mkdtemp(prefix="vgpu_bench_")
    for d in ("etc/vgpu-manager/config", "lock", "vmem"):
        os.makedirs(os.path.join(sandbox, d), exist_ok=True)
    env = tenant_env(H, H.NEW_SO, gpu, uuids, sandbox, 0, extra)
    if have_mountns():
        env["VGPU_TENANT_PRELOAD"] = env.pop("LD_PRELOAD", "")
    else:
        env["SCENARIO_LEDGER"] = os.path.join(sandbox, "vmem", "vmem_node.config")
    r = subprocess.run(in_container([H.SCENARIO], sandbox), env=env, input=script, capture_output=True, text=True, timeout=300,
                       preexec_fn=H.pin_to(H.gpu_local_cpus(gpu)))
    shutil.rmtree(sandbox, ignore_errors=True)
    lines = r.stdout.splitlines()
    st = [l for l in lines if l.startswith("slabstats")]
    if r.returncode != 0 or not st or "none" in st[-1] or "slab mode disabled" in r.stderr:
        return {"error": (r.stderr or r.stdout)[-400:]}
    f = st[-1].split()
    v = dict(zip(f[1::2], map(int, f[2::2])))
    peak = peaks.get("hbm_gbs", 6650.0)
    out = {"workload": "config 4 shape x4 (32 GiB cap oversold x4 => 8 GiB physical, ledger on), VGPU_B200_SLAB=1: 16 x cuMemAlloc(1 GiB) "
                       "+ fill + check through the hooks",
           "intact": sum("intact" in l for l in lines), "corrupt": sum("CORRUPT" in l for l in lines), "counters": v,
           "nvml_view": next((l for l in lines if l.startswith("nvmlinfo")), None),
           "ledger": next((l for l in lines if l.startswith("ledger")), None)}
    if v.get("scrub_ns"):
        g = v["scrubbed_bytes"] / v["scrub_ns"]
        out["scrub"] = {"kernel": "vgpu_clear_kernel (launched by the cuMemAlloc hook)", "bound": "hbm", "achieved": round(g, 1), "unit": "GB/s",
                        "peak": peak, "frac": round(g / peak, 4), "algorithmic_bytes": v["scrubbed_bytes"], "launches": v["demotions"]}
    if v.get("spill_ns"):
        out["spill_to_host"] = {"kernel": "vgpu_spill_copy_kernel (launched by the cuMemAlloc hook)", "bound": "pcie",
                                "achieved": round(v["spill_bytes"] / v["spill_ns"], 2), "unit": "GB/s", "bytes": v["spill_bytes"],
                                "note": "HBM -> host-resident backing under the same virtual address; PCIe Gen5 x16 bound, not HBM"}
    return out


def bandwidth_kernels(lib_path, peaks):
    """vgpu_spill_copy_kernel and vgpu_clear_kernel in-process: 1 GiB buffers (>> 126 MB L2), CUDA
    events on the launching stream, 3 warm-ups, 10 timed launches each."""
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
    del src, dst
    torch.cuda.empty_cache()  # HOST mode counts this process too: give the memory back before the tenants' memory legs
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = 2 * n / avg / 1e9
    return {"bound": "hbm", "kernel": "vgpu_spill_copy_kernel", "achieved": round(achieved, 1), "peak": peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)",
            "unit": "GB/s", "frac": round(achieved / peak, 4),
            # not measured by this run: dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this
            # size (ncu --set full, round 2, profiles/ncu_bandwidth_kernels_r2.csv: 1.073814 + 1.026304 GB)
            "traffic": None, "traffic_ncu_r2_constant": 2100118000,
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
    ap.add_argument("--core-limit", type=int, default=0, help="override the config's core cap (percent)")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-rebalance", action="store_true", help="N > 1: do not run the cross-tenant rebalance loop beside the tenants")
    ap.add_argument("--no-extras", action="store_true", help="skip the bare / allocator / cpu_baseline legs")
    ap.add_argument("--roofline-only", action="store_true",
                    help="only the in-process bandwidth kernels (the form profiled under ncu: a tenant that runs with the "
                         "interposer preloaded cannot also run under ncu's - its gate waits deadlock with kernel serialisation)")
    args = ap.parse_args()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    if args.roofline_only:
        import helpers
        helpers.build_all()
        roof, n = bandwidth_kernels(helpers.NEW_SO, peaks)
        print(json.dumps({"roofline": roof, "gpu_launches": n}))
        return
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    core_limit = args.core_limit or (25 if world == 1 else 50)
    config_name = "configs[1]" if world == 1 else "configs[4]"

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
        if distributed:
            dist.destroy_process_group()
        return
    if args.impl == "b200" and not os.path.exists(H.NEW_SO):
        raise SystemExit("libvgpu-control.so (with its sm_100a image) is missing - no CPU fallback exists")
    uuids = gpu_uuids()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # the reference arm throttles on the CPU in 10 ms sleeps: bound each of its steps
    max_seconds = 0.0
    per_step = args.per_step
    if args.impl == "reference":
        max_seconds = 20.0 * args.steps / 5.0

    # N > 1, B200 arm: every rank also plays its GPU's node agent while the tenant runs - gather -> plan ->
    # apply once per control period (the only collective of the job)
    agent, loops = None, []
    if distributed and args.impl == "b200" and not args.no_rebalance:
        from vgpu_manager_b200.multi import RebalanceLoop
        rounds = int(((args.steps + args.warmup) * per_step / 450e3 + 1.0) / 0.08)

        def agent(sandbox):
            lp = RebalanceLoop(dist, torch, torch.device("cuda", local_rank), local_rank, core_limit,
                               os.path.join(sandbox, "etc/vgpu-manager/config"), os.path.join(sandbox, "lock"), rounds)
            loops.append(lp)
            lp.start()
            return lp

    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    t_region0 = time.perf_counter()
    res = run_tenant(H, lib, local_rank, uuids, args.steps, args.warmup, per_step, core_limit, max_seconds, agent=agent)
    barrier()
    region_s = time.perf_counter() - t_region0
    clk, util_pct = clocks.stop()

    # timed K steps on the device (events inside the tenant); max over ranks
    dev_s = res["device_s"] if res.get("device_s", 0) > 0 else res["wall_s"]
    vals = torch.tensor([dev_s, float(res["launches"]), res["wall_s"], float(res["p50_ns"]), float(res["p99_ns"]),
                         float(res.get("sampler_launches", 0)), float(res.get("gated_launches", 0)),
                         float(util_pct if util_pct is not None else -1), float(res.get("watchdog_loans", 0))],
                        dtype=torch.float64, device="cuda")
    rebalance = None
    if distributed:
        gathered = torch.zeros(world * vals.numel(), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(gathered, vals)
        rows = gathered.view(world, -1).tolist()
        if args.impl == "b200":
            from vgpu_manager_b200.multi import time_collective
            lp = loops[0] if loops else None
            rebalance = {"collective": "all_gather_into_tensor, 4 x f32 per rank, pre-allocated buffers",
                         "collective_us": round(time_collective(dist, torch, torch.device("cuda", local_rank)), 2),
                         "rounds_during_storm": lp.rounds_run if lp else 0, "period_ms": 80,
                         "schedule": "one round per 80 ms while any tenant is under pressure, backing off to 640 ms after 4 calm tables "
                                     "(same decision on every rank from the same gathered table)",
                         "longest_period_ms": round(1e3 * max(lp.periods), 1) if lp and lp.periods else None,
                         "applied_rank0": lp.applied if lp else 0,
                         "targets_rank0_pct": sorted(set(lp.plans)) if lp else []}
    else:
        rows = [vals.tolist()]

    roof, own_launches = None, 0
    cpu_base = bare = alloc_path = None
    extras = rank == 0 and world == 1 and not args.no_extras
    if extras:
        # bare driver, same tenant, no preload: what "added" latency is measured against
        b = run_tenant(H, None, local_rank, uuids, 2, 1, per_step, 0)
        bdev = b["device_s"] if b.get("device_s", 0) > 0 else b["wall_s"]
        bare = {"launches_per_s": round(b["launches"] / bdev, 1), "p50_ns": b["p50_ns"], "p99_ns": b["p99_ns"],
                "sample": "1 warm-up + 2 timed steps of %d launches, no LD_PRELOAD" % per_step}
    mem_path = None
    if extras and args.impl == "b200":
        if not args.no_roofline:
            roof, own_launches = bandwidth_kernels(H.NEW_SO, peaks)
            mem_path = slab_leg(H, local_rank, uuids, peaks)
            if mem_path.get("scrub"):
                # the HBM-bound kernel the memory hooks themselves launched in this run; the in-process 1 GiB copy /
                # clear figures stay alongside as `utility`
                roof = {"bound": "hbm", "kernel": mem_path["scrub"]["kernel"], "achieved": mem_path["scrub"]["achieved"],
                        "peak": mem_path["scrub"]["peak"], "peak_source": roof["peak_source"], "unit": "GB/s",
                        "frac": mem_path["scrub"]["frac"],
                        # not measured by this run: dram__bytes_read.sum + dram__bytes_write.sum of one 1 GiB launch of this kernel
                        # (ncu --set full, round 2, profiles/ncu_bandwidth_kernels_r2.csv: 0.000033 + 1.013717 GB)
                        "traffic": None, "traffic_ncu_r2_constant": 1013750000,
                        "algorithmic_bytes_per_launch": mem_path["scrub"]["algorithmic_bytes"] // max(mem_path["scrub"]["launches"], 1),
                        "launches_timed": mem_path["scrub"]["launches"], "timed_by": "CUDA events recorded by the hook around its own launch",
                        "utility": roof}
                own_launches += 3 * mem_path["counters"].get("demotions", 0) + mem_path["counters"].get("allocs", 0)
        alloc_path = {"workload": "config 4 cap (8 GiB, oversold x4 => 2 GiB physical, ledger on): 1500 x {cuMemAlloc 1 MiB, cuMemFree}",
                      "b200": run_allocstorm(H, H.NEW_SO, local_rank, uuids)}
        if os.path.exists(H.REF_SO):
            alloc_path["reference"] = run_allocstorm(H, H.REF_SO, local_rank, uuids)
            # bounded sample of the same workload under the reference's CPU watcher path
            ref = run_tenant(H, H.REF_SO, local_rank, uuids, 1, 1, per_step, core_limit, max_seconds=20.0)
            rdev = ref["device_s"] if ref.get("device_s", 0) > 0 else ref["wall_s"]
            cpu_base = {"value": round(ref["launches"] / rdev, 1), "unit": "launches/s", "cores": 1, "kind": "reference",
                        "host_cores_on_box": os.cpu_count(),
                        "sample": "1 warm-up + 1 timed step of up to %d launches (20 s cap, %d done) under "
                                  "oracle/_ref/libvgpu-control.so, same %d%%/4GiB tenant env" % (per_step, ref["launches"], core_limit),
                        "p50_hook_ns": ref["p50_ns"], "p99_hook_ns": ref["p99_ns"], "max_hook_ns": ref["max_ns"]}
    elif extras and args.impl == "reference":
        alloc_path = {"workload": "config 4 cap (8 GiB, oversold x4, ledger on): 1500 x {cuMemAlloc 1 MiB, cuMemFree}",
                      "reference": run_allocstorm(H, H.REF_SO, local_rank, uuids)}
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    t_max = max(r[0] for r in rows)
    total = sum(r[1] for r in rows)
    life_max = max(r[2] for r in rows)
    value = total / t_max
    utils = [r[7] for r in rows if r[7] >= 0]
    p50, p99 = max(r[3] for r in rows), max(r[4] for r in rows)
    line = {
        "metric": "intercepted cuLaunchKernel/sec under %d%% core / 4 GiB cap (p50 hook latency alongside)" % core_limit,
        "value": round(value, 1), "unit": "launches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(t_max / max(res["steps"], 1) * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "%s: 1 tenant per B200, %d%% cores / 4 GiB cap, empty-kernel <<<1,1>>> cuLaunchKernel storm, "
                               "%d launches per step, sync per step%s" %
                               (config_name, core_limit, per_step,
                                "" if world == 1 else " (N > 1 runs BASELINE config 5's 50 % cap on every GPU; N = 1 runs configs[1]'s 25 %)"),
                   "per_step_launches": per_step, "core_limit_pct": core_limit, "mem_limit": MEM_LIMIT,
                   "l2_policy": "storm has no data reuse; bandwidth kernels use 1 GiB buffers (> 126 MB L2)",
                   "tenants": world, "impl_library": os.path.relpath(lib, ROOT),
                   "isolation": "mount namespace per tenant" if have_mountns() else "path-redirect shim (unshare -m not permitted)",
                   "tenant_cpus": "NUMA node of the tenant's GPU (both arms, bare leg too)" if H.gpu_local_cpus(local_rank) else "not pinned"},
        "p50_hook_ns": p50, "p99_hook_ns": p99,
        "achieved_util_pct": round(sum(utils) / len(utils), 1) if utils else None,
        "clocks": clk,
        "e2e": {"value": round(total / life_max, 1), "unit": "launches/s",
                "h2d_bytes_per_step": 16 * per_step if args.impl == "b200" else 0,
                "d2h_bytes_per_step": 8 * per_step // 640 if args.impl == "b200" else 0,
                "what": "same K steps on the tenant's HOST clock around launch calls + device sync, through the "
                        "LD_PRELOADed hook; h2d = ticket + sequence words the hook publishes per launch in pinned "
                        "memory (read by the controller over PCIe), d2h = completion markers (one per 640 launches)"},
        "tenant_process_life_s": round(res["life_s"], 3),
        "gpu_launches": int(sum(r[5] for r in rows)) + own_launches,
        "gated_launches": int(sum(r[6] for r in rows)),
        "watchdog_loans": int(sum(r[8] for r in rows)),
        "limiter_rank0": res.get("limiter"),
        "truncated": bool(res.get("truncated", 0)),
        "region_wall_s": round(region_s, 3),
    }
    if bare:
        line["bare"] = bare
        line["added_p50_ns"] = p50 - bare["p50_ns"]
        line["added_p99_ns"] = p99 - bare["p99_ns"]
    if alloc_path:
        line["alloc_path"] = alloc_path
    if mem_path:
        line["mem_path"] = mem_path
    if rebalance is not None:
        line["rebalance"] = rebalance
        line["rebalance_us"] = rebalance.get("collective_us")
    if args.impl == "reference":
        line["impl"] = "reference"
        line["gpu_launches"] = 0
        line["cpu_baseline"] = {"value": line["value"], "unit": "launches/s", "cores": world, "kind": "reference",
                                "host_cores_on_box": os.cpu_count(),
                                "sample": "%d tenant(s) x %d timed steps of up to %d launches (%.0f s cap) under oracle/_ref; each "
                                          "tenant = 1 launching thread + 1 watcher thread" % (world, res["steps"], per_step, max_seconds)}
        line["e2e"] = {"value": line["value"], "unit": "launches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    else:
        line["roofline"] = roof
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))


if __name__ == "__main__":
    main()
