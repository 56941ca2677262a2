"""Tiny driver for ncu captures: a few launches of each bandwidth kernel on 1 GiB buffers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vgpu_manager_b200 import B200Library

torch.zeros(1, device="cuda")
uuid = "GPU-" + str(torch.cuda.get_device_properties(0).uuid)
lib = B200Library(env={"MANAGER_VISIBLE_DEVICES": uuid, "MANAGER_COMPATIBILITY_MODE": "0"})
lib.attach()
n = int(os.environ.get("BYTES", str(1 << 30)))
src = torch.randint(0, 255, (n,), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
s = torch.cuda.current_stream().cuda_stream
for _ in range(int(os.environ.get("ITERS", "5"))):
    lib.spill_copy(dst.data_ptr(), src.data_ptr(), n, s)
    lib.clear(dst.data_ptr(), n, s)
torch.cuda.synchronize()
# reference point: the copy the driver measured MEASURED_PEAKS.json with
a = src.view(torch.bfloat16)
b = dst.view(torch.bfloat16)
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
print("done")
