"""Tiny driver for ncu captures of the latency-bound control kernels: vgpu_refill_kernel (fold of a
full 1024-sample publication + controller) and vgpu_quota_kernel (three 1024-record lists)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as H
from vgpu_manager_b200 import B200Library

torch.zeros(1, device="cuda")
uuid = "GPU-" + str(torch.cuda.get_device_properties(0).uuid)
lib = B200Library(env={"MANAGER_VISIBLE_DEVICES": uuid, "MANAGER_COMPATIBILITY_MODE": "0"})
lib.attach()
lib.limiter_reset(148, 2048, 25, 0, 1, 1)
for n in (2, 1024):
    u = H.UtilReq()
    u.seq, u.status, u.mode, u.n_samples, u.sys_process_num, u.have_container_pids, u.checktime_us = 1, 2, 2, n, 2, 1, 1000
    for i in range(n):
        u.samples[i].pid, u.samples[i].sm, u.samples[i].ts_us = 100 + i, i % 90, 2000
        u.flags[i] = 1 if i % 3 == 0 else 0
    for _ in range(3):
        lib.refill(u)
    q = H.QuotaReq()
    q.kind, q.mode, q.n_compute, q.n_graphics, q.n_vmem, q.total_memory, q.real_memory = 0, 2, n, n, n, 1 << 34, 1 << 33
    for i in range(n):
        q.compute[i].pid, q.compute[i].used_bytes = 100 + i, 1 << 20
        q.graphics[i].pid, q.graphics[i].used_bytes = 100 + 2 * i, 1 << 20
        q.vmem[i].pid, q.vmem[i].used = 100 + i, 4096
        q.cflags[i] = q.gflags[i] = 1
    q.request = 1 << 20
    for _ in range(3):
        lib.quota_eval(q, H.QuotaRes())
print("vslab ops", H.vslab_model_check(lib, ops=int(os.environ.get("VSLAB_OPS", "200"))))
print("done")
