"""Multi-GPU plumbing: one limiter/allocator instance per physical GPU, no data-path collective.

The only exchange between instances is the *rebalance* (SURVEY.md 8e): once per control period
each GPU's node agent contributes {gpu, quota %, utilisation %, gated fraction}; every rank
computes the same plan from the gathered table and applies its own row by writing the tenant's
`rebalance.config` record (`vgpu_b200_set_limits` does the same from C), which the tenant's tick
thread hands to the on-device controller at its next step.  With torch.distributed this is one
`all_gather_into_tensor` of 4 floats per rank on a pre-allocated buffer (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  There is no reference counterpart - its balance policy lives
inside each process (cuda_hook.c:430-465) - hence no parity oracle, only behaviour tests: a gated
tenant's target rises, an un-gated one's does not, nobody exceeds its ceiling.
"""
import os
import struct
import threading
import time
from dataclasses import dataclass
from typing import Dict, List, Sequence

REBALANCE_MAGIC = 0x4C424756  # kernel_abi.h VGPU_REBALANCE_MAGIC
REC = struct.Struct("<IIii")  # vgpu_rebalance_rec_t
STATUS = struct.Struct("<IIiiiiiiqQQQ")  # vgpu_tenant_status_t


@dataclass
class TenantReport:
    gpu: int
    quota_pct: float
    util_pct: float
    gated_frac: float

    def as_vector(self) -> List[float]:
        return [float(self.gpu), float(self.quota_pct), float(self.util_pct), float(self.gated_frac)]

    @staticmethod
    def from_vector(v: Sequence[float]) -> "TenantReport":
        return TenantReport(int(v[0]), float(v[1]), float(v[2]), float(v[3]))


def aggregate(step_seconds: Sequence[float], launches: Sequence[float]):
    """Whole-job throughput the bench contract asks for: all units / max-over-ranks time."""
    t = max(step_seconds)
    return sum(launches) / t if t > 0 else 0.0, t


def rebalance(reports: Sequence[TenantReport], headroom_pct: float = 100.0) -> Dict[int, float]:
    """Utilisation targets: tenants that were never gated keep their quota; the spare share of
    each GPU (100 - quota) is offered to its gated tenant in proportion to how often it was gated,
    capped at `headroom_pct`.  One tenant per GPU, so this is per-GPU bookkeeping every rank can
    compute identically from the gathered table (no second collective needed)."""
    out = {}
    for r in reports:
        spare = max(0.0, 100.0 - r.quota_pct)
        want = r.quota_pct + spare * min(1.0, max(0.0, r.gated_frac))
        out[r.gpu] = min(headroom_pct, want)
    return out


def write_limits(cfg_dir: str, host_index: int, seq: int, up_limit: int, soft_core: int):
    """Python twin of vgpu_b200_set_limits: one 16-byte record of <cfg_dir>/rebalance.config."""
    path = os.path.join(cfg_dir, "rebalance.config")
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    try:
        if os.fstat(fd).st_size < 16 * REC.size:
            os.ftruncate(fd, 16 * REC.size)
        os.pwrite(fd, REC.pack(REBALANCE_MAGIC, seq & 0xFFFFFFFF or 1, int(up_limit), int(soft_core)), host_index * REC.size)
    finally:
        os.close(fd)


def read_status(lock_dir: str, host_index: int):
    """vgpu_tenant_status_t a tenant publishes while a rebalance.config exists; None if absent."""
    try:
        with open(os.path.join(lock_dir, "vgpu_%d.status" % host_index), "rb") as f:
            raw = f.read(STATUS.size)
        if len(raw) < STATUS.size:
            return None
        v = STATUS.unpack(raw)
        if v[0] != REBALANCE_MAGIC:
            return None
        keys = ("magic", "seq", "pid", "user_current", "sys_current", "up_limit", "hard_core", "soft_core", "share",
                "gated", "launched", "steps")
        return dict(zip(keys, v))
    except OSError:
        return None


class AllGather:
    """One pre-allocated all_gather_into_tensor of 4 floats per rank (falls back to the list form on
    backends without it, i.e. gloo in the CPU tests)."""

    def __init__(self, dist, torch, device):
        self.dist, self.torch = dist, torch
        self.world = dist.get_world_size()
        self.vec = torch.zeros(4, dtype=torch.float32, device=device)
        self.out = torch.zeros(4 * self.world, dtype=torch.float32, device=device)
        self.flat = True
        try:
            dist.all_gather_into_tensor(self.out, self.vec)
        except Exception:
            self.flat = False
            self.bufs = [torch.zeros_like(self.vec) for _ in range(self.world)]

    def __call__(self, report: TenantReport) -> List[TenantReport]:
        self.vec.copy_(self.torch.tensor(report.as_vector(), dtype=self.torch.float32), non_blocking=False)
        if self.flat:
            self.dist.all_gather_into_tensor(self.out, self.vec)
            rows = self.out.view(self.world, 4).tolist()
        else:
            self.dist.all_gather(self.bufs, self.vec)
            rows = [b.tolist() for b in self.bufs]
        return [TenantReport.from_vector(r) for r in rows]


def all_gather_reports(dist, torch, report: TenantReport, device):
    """One collective per control period (convenience form)."""
    return AllGather(dist, torch, device)(report)


class RebalanceLoop(threading.Thread):
    """The node agent of one GPU while its tenant runs: read the tenant's status, gather, plan, apply
    the own row, sleep one period.  The collective is the synchronisation, so all ranks must run the
    same number of rounds; the schedule is therefore a pure function of the gathered tables (which
    are bit-identical on every rank): the loop covers `rounds * period_s` of *scheduled* time, one
    control period per round while any tenant of the job is under pressure, and - every rank taking
    the same decision from the same table - twice the previous period, up to `max_backoff` periods,
    once `calm_rounds` consecutive tables showed none.  A collective's kernel is another process's
    work on the tenant's time-sliced GPU; an un-throttled job should not pay for 12 of them a second."""

    def __init__(self, dist, torch, device, gpu, quota_pct, cfg_dir, lock_dir, rounds, period_s=0.08, ceiling=100,
                 host_index=None, max_backoff=8, calm_rounds=4):
        """`gpu` is this agent's key in the gathered table (the physical GPU on a real node, the rank in the
        CPU tests); `host_index` the tenant's index of that GPU inside its own config (defaults to `gpu`).
        `max_backoff=1` keeps the fixed cadence."""
        super().__init__(daemon=True)
        self.gather = AllGather(dist, torch, device)
        self.torch, self.device = torch, device
        self.gpu, self.quota, self.cfg_dir, self.lock_dir = gpu, quota_pct, cfg_dir, lock_dir
        self.host_index = gpu if host_index is None else host_index
        self.pressure = 0.0
        self.rounds, self.period_s, self.ceiling = rounds, period_s, ceiling
        self.max_backoff, self.calm_rounds = max(1, int(max_backoff)), max(1, int(calm_rounds))
        self.rounds_run, self.periods = 0, []
        self.applied, self.plans, self.seq = 0, [], 0
        self.tenant_up_limits = set()  # what the tenant's controller reported back (status file)
        self.last = None
        # the tenant publishes its status only once it has seen a rebalance.config
        write_limits(cfg_dir, self.host_index, 0, 0, 0)

    def run(self):
        prev = None
        if str(self.device).startswith("cuda"):
            self.torch.cuda.set_device(self.device)  # the current device is per thread
        scheduled, total = 0.0, self.rounds * self.period_s
        period, calm = self.period_s, 0
        while scheduled < total - 1e-9:
            t0 = time.perf_counter()
            st = read_status(self.lock_dir, self.host_index)
            util = 0.0
            hit = 0.0
            if st:
                util = float(st["user_current"])
                self.tenant_up_limits.add(int(st["up_limit"]))
                if prev and st["gated"] > prev["gated"]:
                    hit = 1.0  # the tenant ran into its cap during this period
                prev = st
            # "how hard is the cap biting": smoothed over a few periods so that the target moves gradually
            self.pressure = 0.7 * self.pressure + 0.3 * hit
            table = self.gather(TenantReport(self.gpu, self.quota, util, self.pressure))
            plan = rebalance(table, self.ceiling)
            mine = int(round(plan[self.gpu] / 5.0)) * 5  # 5 % steps: the controller's own minimum step
            if mine != self.last:
                self.seq += 1
                # a target above the quota needs a ceiling above it too (balance mode); back at the quota
                # the ceiling goes back as well and the tenant returns to its hard limit
                write_limits(self.cfg_dir, self.host_index, self.seq, mine, self.ceiling if mine > self.quota else 0)
                self.last = mine
                self.applied += 1
            self.plans.append(mine)
            # next period: from the gathered table only, so that every rank schedules the same rounds
            if any(r.gated_frac > 0.01 for r in table):
                period, calm = self.period_s, 0
            else:
                calm += 1
                if calm >= self.calm_rounds:
                    period = min(period * 2.0, self.period_s * self.max_backoff)
            self.rounds_run += 1
            self.periods.append(period)
            scheduled += period
            rest = period - (time.perf_counter() - t0)
            if rest > 0:
                time.sleep(rest)


def time_collective(dist, torch, device, iters=50):
    """Device time of the rebalance collective itself (CUDA events, max over ranks), microseconds."""
    ag = AllGather(dist, torch, device)
    rep = TenantReport(0, 50.0, 10.0, 0.0)
    for _ in range(10):
        ag(rep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        if ag.flat:
            dist.all_gather_into_tensor(ag.out, ag.vec)
        else:
            dist.all_gather(ag.bufs, ag.vec)
    e1.record()
    e1.synchronize()
    us = torch.tensor([e0.elapsed_time(e1) * 1e3 / iters], dtype=torch.float64, device=device)
    dist.all_reduce(us, op=dist.ReduceOp.MAX)
    return float(us.item())
