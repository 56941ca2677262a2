"""Multi-GPU plumbing: one limiter/allocator instance per physical GPU, no data-path collective.

The only exchange between instances is the *rebalance vector* (SURVEY.md 8e): every control
period each GPU's tenant contributes {gpu, quota %, achieved launches/s, gated fraction}; the
gathered table is what a node-level policy (the reference's `balance` compute policy lives inside
each process, cuda_hook.c:430-465) would redistribute soft limits from.  With torch.distributed
this is one all_gather of 4 floats per rank (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  There is no reference counterpart, hence no parity oracle - only shape/aggregation tests.
"""
from dataclasses import dataclass
from typing import List, Sequence


@dataclass
class TenantReport:
    gpu: int
    quota_pct: float
    achieved_per_s: float
    gated_frac: float

    def as_vector(self) -> List[float]:
        return [float(self.gpu), float(self.quota_pct), float(self.achieved_per_s), float(self.gated_frac)]

    @staticmethod
    def from_vector(v: Sequence[float]) -> "TenantReport":
        return TenantReport(int(v[0]), float(v[1]), float(v[2]), float(v[3]))


def aggregate(step_seconds: Sequence[float], launches: Sequence[float]):
    """Whole-job throughput the bench contract asks for: all units / max-over-ranks time."""
    t = max(step_seconds)
    return sum(launches) / t if t > 0 else 0.0, t


def rebalance(reports: Sequence[TenantReport], headroom_pct: float = 100.0):
    """Proposed soft limits: tenants that were never gated keep their quota; the spare share of
    each GPU (100 - quota) is offered to its gated tenant, capped at `headroom_pct`.
    One tenant per GPU, so this is per-GPU bookkeeping that every rank can compute identically
    from the gathered table (no second collective needed)."""
    out = {}
    for r in reports:
        spare = max(0.0, 100.0 - r.quota_pct)
        want = r.quota_pct + spare * min(1.0, r.gated_frac)
        out[r.gpu] = min(headroom_pct, want)
    return out


def all_gather_reports(dist, torch, report: TenantReport, device):
    """One collective per control period."""
    vec = torch.tensor(report.as_vector(), dtype=torch.float32, device=device)
    bufs = [torch.zeros_like(vec) for _ in range(dist.get_world_size())]
    dist.all_gather(bufs, vec)
    return [TenantReport.from_vector(b.tolist()) for b in bufs]
