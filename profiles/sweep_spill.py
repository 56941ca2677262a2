"""Sweep the spill-copy geometry on a B200 (chunk x stages x CTAs/SM), 1 GiB HBM->HBM, CUDA events."""
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
n = 1 << 30
src = torch.randint(0, 255, (n,), dtype=torch.uint8, device="cuda")
dst = torch.empty_like(src)
stream = torch.cuda.current_stream()


def timeit(fn, iters=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return sum(ts) / len(ts), min(ts)


rows = []
a16, b16 = src.view(torch.bfloat16), dst.view(torch.bfloat16)
avg, best = timeit(lambda: b16.copy_(a16))
rows.append({"kernel": "torch copy_ (MEASURED_PEAKS method)", "avg_gbs": 2 * n / avg / 1e9, "best_gbs": 2 * n / best / 1e9})
avg, best = timeit(lambda: lib.clear(dst.data_ptr(), n, stream.cuda_stream))
rows.append({"kernel": "vgpu_clear", "avg_gbs": n / avg / 1e9, "best_gbs": n / best / 1e9})
for chunk in (8192, 16384, 32768, 65536):
    for stages in (2, 3, 4, 6, 8, 12):
        for ctas in (1, 2, 3, 4, 6, 8):
            if chunk * stages > 200 * 1024 or chunk * stages * ctas > 227 * 1024:
                continue
            lib.set_spill_geometry(chunk, stages, ctas)
            avg, best = timeit(lambda: lib.spill_copy(dst.data_ptr(), src.data_ptr(), n, stream.cuda_stream), iters=5)
            rows.append({"kernel": "vgpu_spill_copy", "chunk": chunk, "stages": stages, "ctas_per_sm": ctas,
                         "avg_gbs": round(2 * n / avg / 1e9, 1), "best_gbs": round(2 * n / best / 1e9, 1)})
assert torch.equal(src, dst)
rows_sorted = sorted([r for r in rows if r["kernel"] == "vgpu_spill_copy"], key=lambda r: -r["avg_gbs"])
print(json.dumps({"reference_rows": rows[:2], "top": rows_sorted[:12], "all": rows_sorted}))
