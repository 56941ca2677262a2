"""GPU x2: a distributed PyTorch job (torchrun, NCCL all_reduce + bf16 GEMMs) as the tenant, one rank
per GPU, both ranks in one "container" (shared vgpu.config / lock files), under each library.
NCCL exercises what a single-GPU tenant does not: cuMemCreate/cuMemMap (VMM) buffers, CUDA IPC,
several internal streams, cuLaunchKernelEx with launch attributes.  The job must produce the same
numbers and see the capped memory on both devices."""
import json
import os
import subprocess
import sys

import pytest

import helpers as H
from test_gpu_differential import gpu_uuids

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(len(gpu_uuids()) < 2, reason="needs two GPUs (gpurun --gpus 2)")]

JOB = r'''
import json, os, torch, torch.distributed as dist
dist.init_process_group("nccl")
r = dist.get_rank()
torch.cuda.set_device(r)
x = torch.full((1 << 20,), float(r + 1), device="cuda")
for _ in range(20):
    dist.all_reduce(x)
    x /= 2.0
a = torch.full((2048, 2048), 0.001, device="cuda", dtype=torch.bfloat16)
for _ in range(100):
    b = a @ a
torch.cuda.synchronize()
free, total = torch.cuda.mem_get_info()
out = torch.tensor([float(total), float(x[0]), float(b[0, 0])], device="cuda", dtype=torch.float64)
gathered = [torch.zeros_like(out) for _ in range(dist.get_world_size())]
dist.all_gather(gathered, out)
if r == 0:
    print(json.dumps([g.tolist() for g in gathered]))
dist.barrier()
dist.destroy_process_group()
'''


def run_job(lib, port):
    sb = H.Sandbox()
    u = gpu_uuids()
    knobs = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": ",".join(u[:2]), "CUDA_VISIBLE_DEVICES": "0,1",
             "CUDA_MEM_LIMIT_0": "16g", "CUDA_MEM_LIMIT_1": "12g", "CUDA_CORE_LIMIT_0": "50", "CUDA_CORE_LIMIT_1": "50",
             "LOGGER_LEVEL": "1"}
    env = H.preload_env(lib, sb, knobs, stub=False) if lib else dict(os.environ, CUDA_VISIBLE_DEVICES="0,1")
    script = os.path.join(sb.dir, "job.py")
    with open(script, "w") as f:
        f.write(JOB)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), script], env=env, capture_output=True, text=True, timeout=400)
    sb.cleanup()
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("[[")][-1])


def test_two_rank_nccl_job_under_both_libraries(built):
    bare = run_job(None, 29611)
    ours = run_job(H.NEW_SO, 29612)
    assert [row[1:] for row in ours] == [row[1:] for row in bare]  # all_reduce chain and GEMM values
    assert ours[0][0] == float(16 << 30) and ours[1][0] == float(12 << 30)  # each rank sees its device's cap
    report = {"bare": bare, "b200": ours}
    if os.path.exists(H.REF_SO):
        ref = run_job(H.REF_SO, 29613)
        report["reference"] = ref
        assert ref == ours
    os.makedirs(os.path.join(H.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(H.ROOT, "gpurun_out", "nccl_tenant_r1.json"), "w") as f:
        json.dump(report, f, indent=1)
