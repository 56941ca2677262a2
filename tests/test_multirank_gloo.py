"""CPU: the N>1 path of bench.py with world_size 2 over gloo - one stub-driver tenant per rank,
independent limiter instances, the rebalance all_gather, whole-job aggregation."""
import json
import os
import subprocess
import sys

import helpers as H


def free_port():
    """A rendezvous port nobody listens on right now (a fixed one collides with a second suite on the same machine)."""
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        return str(so.getsockname()[1])


WORKER = r'''
import json, os, subprocess, sys
sys.path.insert(0, os.environ["REPO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
import torch, torch.distributed as dist
import helpers as H
from vgpu_manager_b200.multi import TenantReport, aggregate, rebalance, all_gather_reports
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sb = H.Sandbox()
env = H.preload_env(H.NEW_SO, sb, {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": H.STUB_UUID,
                                  "CUDA_CORE_LIMIT_0": "50", "CUDA_MEM_LIMIT_0": "1g", "STUB_UTIL": "closed:0.02"})
dist.barrier()
r = subprocess.run([H.STORM, "--steps", "2", "--warmup", "1", "--per-step", "50000", "--no-kernel"], env=env,
                   capture_output=True, text=True, timeout=120)
dist.barrier()
d = json.loads(r.stdout.strip().splitlines()[-1])
rep = TenantReport(rank, 50.0, 10.0, d["gated_launches"] / max(d["launches"], 1))
table = all_gather_reports(dist, torch, rep, "cpu")
t = torch.tensor([d["wall_s"], float(d["launches"])], dtype=torch.float64)
rows = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(rows, t)
value, tmax = aggregate([float(x[0]) for x in rows], [float(x[1]) for x in rows])
if rank == 0:
    print(json.dumps({"value": value, "tmax": tmax, "launches": sum(float(x[1]) for x in rows),
                      "gpus": [r.gpu for r in table], "plan": rebalance(table), "sampler": d["sampler_launches"]}))
sb.cleanup()
dist.destroy_process_group()
'''


def test_two_ranks_independent_tenants_and_rebalance_gather(built, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=H.ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", free_port(), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["launches"] == 200000 and d["gpus"] == [0, 1]
    assert abs(d["value"] - d["launches"] / d["tmax"]) < 1e-6
    assert set(d["plan"].keys()) == {"0", "1"} and all(50.0 <= v <= 100.0 for v in d["plan"].values())


REBALANCE_WORKER = r'''
import json, os, subprocess, sys
sys.path.insert(0, os.environ["REPO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
import torch, torch.distributed as dist
import helpers as H
from vgpu_manager_b200.multi import RebalanceLoop
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sb = H.Sandbox()
env = H.preload_env(H.NEW_SO, sb, {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": H.STUB_UUID,
                                  "CUDA_CORE_LIMIT_0": "50", "CUDA_MEM_LIMIT_0": "1g", "LOGGER_LEVEL": "3",
                                  "STUB_UTIL": "closed:0.02" if rank == 0 else "fixed:5"})
loop = RebalanceLoop(dist, torch, "cpu", rank, 50, sb.path("etc/vgpu-manager/config"), sb.path("lock"), rounds=80, period_s=0.08,
                     host_index=0)
dist.barrier()
loop.start()
if rank == 0:   # saturates its cap: gated
    cmd = [H.STORM, "--steps", "100000", "--warmup", "0", "--per-step", "40000", "--no-kernel", "--max-seconds", "5"]
else:           # a trickle of launches: never gated
    cmd = [H.SCENARIO]
r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=120,
                   input=None if rank == 0 else "init 0\n" + "launch 20 1 1 1\nsleepms 300\n" * 16)
loop.join(timeout=60)
out = {"rank": rank, "rc": r.returncode, "applied": loop.applied, "plans": sorted(set(loop.plans)),
       "tenant_up_limits": sorted(loop.tenant_up_limits), "assigned_log": r.stderr.count("node agent assigned"),
       "rounds_run": loop.rounds_run, "periods": [round(p, 3) for p in loop.periods]}
rows = [None] * world
dist.all_gather_object(rows, out)
if rank == 0:
    print(json.dumps(rows))
sb.cleanup()
dist.destroy_process_group()
'''


def test_rebalance_is_applied_to_the_gated_tenant_only(built, tmp_path):
    """gather -> plan -> apply once per control period while both tenants run: the rank whose tenant is
    gated gets a target above its quota (written to rebalance.config, picked up by the tenant's tick
    thread, applied by the on-device controller and reported back through the status file); the
    rank whose tenant never hits its cap keeps exactly its quota."""
    script = tmp_path / "worker.py"
    script.write_text(REBALANCE_WORKER)
    env = dict(os.environ, REPO_ROOT=H.ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", free_port(), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    rows = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    gated, calm = rows[0], rows[1]
    assert gated["rc"] == 0 and calm["rc"] == 0, rows
    assert max(gated["plans"]) > 50 and gated["applied"] >= 2, gated
    assert max(gated["tenant_up_limits"]) > 50, gated       # the controller in HBM really took the new target
    assert gated["assigned_log"] >= 1, gated
    assert calm["plans"] == [50] and calm["applied"] == 1, calm
    assert all(u <= 50 for u in calm["tenant_up_limits"]), calm
    # the schedule is a function of the gathered table only: both ranks ran the same rounds at the same periods
    assert gated["rounds_run"] == calm["rounds_run"] and gated["periods"] == calm["periods"], rows
    assert min(gated["periods"]) == 0.08


CALM_WORKER = r'''
import json, os, sys
sys.path.insert(0, os.environ["REPO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["REPO_ROOT"], "tests"))
import torch, torch.distributed as dist
import helpers as H
from vgpu_manager_b200.multi import RebalanceLoop
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sb = H.Sandbox()
os.makedirs(sb.path("etc/vgpu-manager/config"), exist_ok=True)
loop = RebalanceLoop(dist, torch, "cpu", rank, 50, sb.path("etc/vgpu-manager/config"), sb.path("lock"), rounds=50, period_s=0.02, host_index=0)
dist.barrier()
loop.start()
loop.join(timeout=60)
rows = [None] * world
dist.all_gather_object(rows, {"rounds_run": loop.rounds_run, "periods": [round(p, 3) for p in loop.periods], "plans": sorted(set(loop.plans))})
if rank == 0:
    print(json.dumps(rows))
sb.cleanup()
dist.destroy_process_group()
'''


def test_rebalance_loop_backs_off_identically_on_every_rank_when_nobody_is_gated(built, tmp_path):
    """No tenant under pressure: after four calm tables the period doubles up to 8 control periods, every
    rank taking the same decision from the same gathered table - so the collective count stays matched
    (the loops end together) while an un-throttled job pays for far fewer collectives."""
    script = tmp_path / "calm.py"
    script.write_text(CALM_WORKER)
    env = dict(os.environ, REPO_ROOT=H.ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", free_port(), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    rows = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    assert rows[0] == rows[1], rows
    assert rows[0]["rounds_run"] < 20 and max(rows[0]["periods"]) == 0.16 and rows[0]["plans"] == [50], rows


def test_rebalance_policy_shapes():
    from vgpu_manager_b200.multi import TenantReport, aggregate, rebalance
    plan = rebalance([TenantReport(0, 25, 1e5, 0.0), TenantReport(1, 25, 1e5, 1.0), TenantReport(2, 50, 1e5, 0.5)])
    assert plan == {0: 25.0, 1: 100.0, 2: 75.0}
    assert aggregate([1.0, 2.0], [10, 30]) == (20.0, 2.0)
