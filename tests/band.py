"""Enforcement shapes and the tolerance band derived from the reference itself.

The reference defines no numeric core-% tolerance (SURVEY.md 8a L-tol), north_star asks for "the
reference's own tolerance": so the band is PRODUCED - `python tests/band.py --impl reference --runs 5`
on the GPU box runs the unmodified reference library (oracle/_ref) through every shape below and
writes {min, max, values} per metric; the committed result is tests/golden/tolerance_band.json.
tests/test_gpu_band.py then runs the B200 library through the same shapes and asserts every metric
inside [min - spread, max + spread], spread = max(max - min, 3 % of the mean).

Shapes (all on GPU 0, HOST compatibility mode, tenants = tests/harness/storm.c or a torch GEMM loop):
  storm10 / storm25 / storm50   one tenant, empty-kernel <<<1,1>>> storm under a 10 / 25 / 50 % cap
                                (BASELINE configs[1], config 5's per-GPU load)
  neighbour                     tenant A capped at 10 % saturating its cap with busy kernels, tenant B beside it
                                without any library: A's rate, B's rate relative to B alone
  fair4                         four tenants x 25 %, busy kernels: per-tenant rate, max/min
  gemm1 / gemm4                 BASELINE config 3: one / four tenants x 25 %, bf16 4096^3 GEMM loop: share of
                                an un-capped tenant's rate
TEST / BENCH INFRASTRUCTURE.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import helpers as H  # noqa: E402

BAND_FILE = os.path.join(HERE, "golden", "tolerance_band.json")
BUSY = ["--spin-iters", "20000", "--grid", "592", "--block", "256"]

GEMM_TENANT = r'''
import torch, time, json, os
n = 4096
a = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
b = torch.randn(n, n, dtype=torch.bfloat16, device="cuda")
for _ in range(10):
    c = a @ b
torch.cuda.synchronize()
t0 = time.time()
done = 0
while time.time() - t0 < float(os.environ["TENANT_SECONDS"]):
    for _ in range(25):
        c = a @ b
    torch.cuda.synchronize()
    done += 25
wall = time.time() - t0
print(json.dumps({"gemms": done, "wall_s": wall, "gemms_per_s": done / wall, "pid": os.getpid()}))
'''


_CPUS = "unset"


def local_cpus():
    global _CPUS
    if _CPUS == "unset":
        _CPUS = H.gpu_local_cpus(0)
    return _CPUS


def gpu0_uuid():
    out = subprocess.run(["nvidia-smi", "--query-gpu=uuid", "--format=csv,noheader"], capture_output=True, text=True)
    return out.stdout.splitlines()[0].strip()


class UtilSampler(threading.Thread):
    """nvidia-smi utilization.gpu of GPU 0 every 200 ms (what an operator would look at)."""

    def __init__(self):
        super().__init__(daemon=True)
        self.vals, self.proc = [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=utilization.gpu", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                try:
                    self.vals.append(float(line.strip()))
                except ValueError:
                    pass
        except Exception:
            pass

    def stop(self, skip_s=2.0):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        v = self.vals[int(skip_s / 0.2):] or self.vals
        return round(sum(v) / len(v), 2) if v else None


def tenant_env(lib, sb, cap, mem="4g"):
    knobs = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": gpu0_uuid(), "CUDA_VISIBLE_DEVICES": "0",
             "CUDA_MEM_LIMIT_0": mem, "LOGGER_LEVEL": os.environ.get("BAND_LOGGER_LEVEL", "1")}
    if cap:
        knobs["CUDA_CORE_LIMIT_0"] = str(cap)
    if lib:
        return H.preload_env(lib, sb, knobs, stub=False)
    return dict(os.environ, CUDA_VISIBLE_DEVICES="0")


def start_storm(lib, cap, seconds, busy=False):
    sb = H.Sandbox()
    cmd = [H.STORM, "--steps", "1000000", "--warmup", "0", "--per-step", "200" if busy else "200000", "--max-seconds", str(seconds)]
    if busy:
        cmd += BUSY
    p = subprocess.Popen(cmd, env=tenant_env(lib, sb, cap), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         preexec_fn=H.pin_to(local_cpus()))
    return p, sb


_DETAIL = []  # BAND_DETAIL=<file>: every tenant's counters and stderr tail are appended there (diagnostics)


def finish(p, sb, seconds):
    out, err = p.communicate(timeout=seconds * 8 + 120)
    sb.cleanup()
    if p.returncode != 0 or not out.strip():
        raise RuntimeError("tenant failed rc=%s\n%s" % (p.returncode, err[-2000:]))
    d = json.loads(out.strip().splitlines()[-1])
    d["rate"] = d["launches"] / d["wall_s"]
    if os.environ.get("BAND_STDERR_DIR"):
        os.makedirs(os.environ["BAND_STDERR_DIR"], exist_ok=True)
        with open(os.path.join(os.environ["BAND_STDERR_DIR"], "tenant_%d.err" % p.pid), "w") as f:
            f.write("# rate %.1f\n" % d["rate"] + err)
    if os.environ.get("BAND_DETAIL"):
        _DETAIL.append({"pid": p.pid, "rate": d["rate"], "gated": d.get("gated_launches"), "loans": d.get("watchdog_loans"),
                        "refills": d.get("sampler_launches"), "limiter": d.get("limiter"), "max_ns": d.get("max_ns"),
                        "stderr": err[-6000:]})
        with open(os.environ["BAND_DETAIL"], "w") as f:
            json.dump(_DETAIL, f, indent=1)
    return d


# ----------------------------------------------------------------------------- shapes
def shape_storm(lib, cap, seconds=12.0):
    mon = UtilSampler()
    p, sb = start_storm(lib, cap, seconds)
    mon.start()
    d = finish(p, sb, seconds)
    return {"launches_per_s": d["rate"], "util_pct": mon.stop(), "p50_hook_ns": d["p50_ns"], "p99_hook_ns": d["p99_ns"]}


def neighbour_alone(seconds=8.0):
    p, sb = start_storm(None, 0, seconds, busy=True)
    return finish(p, sb, seconds)["rate"]


def shape_neighbour(lib, alone_rate, seconds=8.0):
    pa, sa = start_storm(lib, 10, seconds + 2.0, busy=True)
    time.sleep(1.0)  # A reaches its throttled regime first
    pb, sbb = start_storm(None, 0, seconds, busy=True)
    b = finish(pb, sbb, seconds)
    a = finish(pa, sa, seconds + 2.0)
    return {"capped_tenant_per_s": a["rate"], "neighbour_vs_alone": b["rate"] / alone_rate}


def shape_fair4(lib, seconds=8.0):
    ps = [start_storm(lib, 25, seconds, busy=True) for _ in range(4)]
    rates = [finish(p, sb, seconds)["rate"] for p, sb in ps]
    return {"tenant_min_per_s": min(rates), "tenant_max_per_s": max(rates), "max_over_min": max(rates) / max(min(rates), 1e-9)}


def run_gemm(lib, count, cap, seconds=12.0):
    procs = []
    for _ in range(count):
        sb = H.Sandbox()
        env = tenant_env(lib, sb, cap, mem="8g")
        env["TENANT_SECONDS"] = str(seconds)
        procs.append((subprocess.Popen([sys.executable, "-c", GEMM_TENANT], env=env, stdout=subprocess.PIPE,
                                       stderr=subprocess.PIPE, text=True, preexec_fn=H.pin_to(local_cpus())), sb))
    rates = []
    for p, sb in procs:
        out, err = p.communicate(timeout=seconds * 8 + 180)
        sb.cleanup()
        lines = [l for l in out.splitlines() if l.startswith("{")]
        if p.returncode != 0 or not lines:
            raise RuntimeError("gemm tenant failed rc=%s\n%s" % (p.returncode, err[-2000:]))
        rates.append(json.loads(lines[-1])["gemms_per_s"])
        if os.environ.get("BAND_STDERR_DIR"):
            os.makedirs(os.environ["BAND_STDERR_DIR"], exist_ok=True)
            with open(os.path.join(os.environ["BAND_STDERR_DIR"], "gemm_%d.err" % p.pid), "w") as f:
                f.write("# gemms_per_s %.1f\n" % rates[-1] + err)
    return rates


def shape_gemm1(lib, alone):
    return {"share_of_alone": run_gemm(lib, 1, 25)[0] / alone}


def shape_gemm4(lib, alone):
    r = run_gemm(lib, 4, 25)
    return {"share_min": min(r) / alone, "share_max": max(r) / alone, "share_sum": sum(r) / alone}


CHEAP = ("storm10", "storm25", "storm50", "neighbour", "fair4")
ALL = CHEAP + ("gemm1", "gemm4")


SETTLE_S = 2.5


def run_shape(name, lib, ctx):
    """ctx carries the un-capped baselines (measured once per session, without any library).
    Every run starts on a GPU that has been idle for SETTLE_S: NVML keeps the previous measurement's
    per-process samples "fresh" for a second, HOST mode sums them into everybody's first readings, and
    the first readings decide how full the bucket gets before the cap starts to bite (the start-up lottery
    visible in both libraries' logs, profiles/README.md)."""
    if name in ("gemm1", "gemm4") and "gemm_alone" not in ctx:
        ctx["gemm_alone"] = run_gemm(None, 1, 0)[0]
    if name == "neighbour" and "neighbour_alone" not in ctx:
        ctx["neighbour_alone"] = neighbour_alone()
    time.sleep(SETTLE_S)
    if name == "storm10":
        return shape_storm(lib, 10)
    if name == "storm25":
        return shape_storm(lib, 25)
    if name == "storm50":
        return shape_storm(lib, 50, 10.0)
    if name == "neighbour":
        if "neighbour_alone" not in ctx:
            ctx["neighbour_alone"] = neighbour_alone()
        return shape_neighbour(lib, ctx["neighbour_alone"])
    if name == "fair4":
        return shape_fair4(lib)
    if name in ("gemm1", "gemm4"):
        if "gemm_alone" not in ctx:
            ctx["gemm_alone"] = run_gemm(None, 1, 0)[0]
        return shape_gemm1(lib, ctx["gemm_alone"]) if name == "gemm1" else shape_gemm4(lib, ctx["gemm_alone"])
    raise KeyError(name)


# ----------------------------------------------------------------------------- band arithmetic
def fold(runs):
    """[{metric: value}] -> {metric: {min, max, mean, values}}"""
    out = {}
    for m in runs[0]:
        vals = [r[m] for r in runs if r.get(m) is not None]
        if vals:
            out[m] = {"min": min(vals), "max": max(vals), "mean": sum(vals) / len(vals), "values": vals}
    return out


def limits(entry, rel_floor=0.03):
    spread = max(entry["max"] - entry["min"], rel_floor * abs(entry["mean"]))
    return entry["min"] - spread, entry["max"] + spread


def check(shape, got, band):
    """-> list of violations (strings); latency metrics are informative, not enforcement, and are not checked"""
    bad = []
    for m, e in band["shapes"][shape].items():
        if m.endswith("_ns") or got.get(m) is None:
            continue
        lo, hi = limits(e)
        if not (lo <= got[m] <= hi):
            bad.append("%s.%s = %.4g outside [%.4g, %.4g] (reference runs: %s)" %
                       (shape, m, got[m], lo, hi, ", ".join("%.4g" % v for v in e["values"])))
    return bad


def load_band():
    with open(BAND_FILE) as f:
        return json.load(f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="reference", choices=["reference", "b200"])
    ap.add_argument("--runs", type=int, default=5)
    ap.add_argument("--shapes", default=",".join(ALL))
    ap.add_argument("--out", default=os.path.join(H.ROOT, "gpurun_out", "tolerance_band.json"))
    ap.add_argument("--append", default=None, help="an existing band file of the same impl: its runs are kept and these are added")
    args = ap.parse_args()
    H.build_all()
    lib = H.REF_SO if args.impl == "reference" else H.NEW_SO
    ctx, shapes = {}, {}
    t0 = time.time()
    for name in args.shapes.split(","):
        runs = []
        for i in range(args.runs):
            r = run_shape(name, lib, ctx)
            runs.append(r)
            print("%s[%d] %s" % (name, i, json.dumps(r)), file=sys.stderr, flush=True)
        shapes[name] = fold(runs)
    runs_note = {args.shapes: args.runs}
    if args.append and os.path.exists(args.append):
        with open(args.append) as f:
            old = json.load(f)
        assert old.get("impl") == args.impl, "cannot mix libraries in one band"
        for name, metrics in old["shapes"].items():
            if name not in shapes:
                shapes[name] = metrics
                continue
            for m, e in metrics.items():
                vals = e["values"] + shapes[name].get(m, {"values": []})["values"]
                shapes[name][m] = {"min": min(vals), "max": max(vals), "mean": sum(vals) / len(vals), "values": vals}
        prev = old.get("runs")
        runs_note = {"earlier sessions": prev, "this session": runs_note}
        for k, v in old.get("baselines", {}).items():
            ctx.setdefault(k + "_earlier", v)
    out = {"impl": args.impl, "library": os.path.relpath(lib, H.ROOT), "runs": runs_note, "baselines": ctx,
           "seconds": round(time.time() - t0, 1), "host_cores": os.cpu_count(),
           "tenant_cpus": "NUMA-local to GPU 0 (%d cpus)" % len(local_cpus()) if local_cpus() else "not pinned",
           "gpu": subprocess.run(["nvidia-smi", "--query-gpu=name,driver_version", "--format=csv,noheader"],
                                 capture_output=True, text=True).stdout.strip(),
           "shapes": shapes}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: {m: [round(e["min"], 4), round(e["max"], 4)] for m, e in v.items()} for k, v in shapes.items()}))


if __name__ == "__main__":
    main()
