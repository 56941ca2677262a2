"""Spill to host-resident pages ('HBM -> host-resident managed = N HBM-read, bounded by
PCIe, report separately'): the TMA spill kernel writing into pinned, mapped host memory and reading it
back, against cudaMemcpy on the same buffers.  CUDA events, 3 warm-ups, 8 timed launches, 256 MiB."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vgpu_manager_b200 import B200Library

torch.zeros(1, device="cuda")
uuid = "GPU-" + str(torch.cuda.get_device_properties(0).uuid)
lib = B200Library(env={"MANAGER_VISIBLE_DEVICES": uuid, "MANAGER_COMPATIBILITY_MODE": "0"})
lib.attach()
n = 256 << 20
dev = torch.randint(0, 255, (n,), dtype=torch.uint8, device="cuda")
back = torch.empty_like(dev)
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
stream = torch.cuda.current_stream()
s = stream.cuda_stream


def timed(fn, reps=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return sum(ts) / len(ts)


out = {"bytes": n}
t = timed(lambda: lib.spill_copy(host.data_ptr(), dev.data_ptr(), n, s))
out["spill_kernel_d2h_gbs"] = round(n / t / 1e9, 2)
assert torch.equal(host, dev.cpu())
t = timed(lambda: lib.spill_copy(back.data_ptr(), host.data_ptr(), n, s))
out["spill_kernel_h2d_gbs"] = round(n / t / 1e9, 2)
assert torch.equal(back, dev)
t = timed(lambda: host.copy_(dev, non_blocking=True))
out["cudaMemcpyAsync_d2h_gbs"] = round(n / t / 1e9, 2)
t = timed(lambda: back.copy_(host, non_blocking=True))
out["cudaMemcpyAsync_h2d_gbs"] = round(n / t / 1e9, 2)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/spill_to_host_r1.json", "w"), indent=1)
print(json.dumps(out))
