"""GPU (B200): vgpu-smwatcher and the exporter against the real NVML - a live CUDA process must show
up in the published process list with its memory, and in the exporter's per-container sample."""
import os
import struct
import subprocess
import sys
import time

import pytest

import helpers as H
from vgpu_manager_b200 import exporter as E

pytestmark = pytest.mark.gpu
WATCHER = os.path.join(H.ROOT, "vgpu_manager_b200", "vgpu-smwatcher")
HOLDER = r'''
import torch, time, sys
x = torch.ones(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(40):
    x.add_(1)
    torch.cuda.synchronize()
    time.sleep(0.1)
'''


def test_live_process_is_published_and_exported(built, tmp_path):
    holder = subprocess.Popen([sys.executable, "-c", HOLDER], env=dict(os.environ, CUDA_VISIBLE_DEVICES="0"))
    try:
        time.sleep(2.5)  # context up, buffer allocated
        path = str(tmp_path / "watcher" / "sm_util.config")
        r = subprocess.run([WATCHER, "--file", path, "--passes", "8", "--period-ms", "150"], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        raw = open(path, "rb").read()
        assert len(raw) == 1311232
        nc = struct.unpack_from("<I", raw, 57360)[0]
        procs = {struct.unpack_from("<IxxxxQII", raw, 32784 + 24 * i)[0]: struct.unpack_from("<IxxxxQII", raw, 32784 + 24 * i)[1]
                 for i in range(nc)}
        assert holder.pid in procs and procs[holder.pid] >= 256 << 20, (holder.pid, procs)
        last_seen = struct.unpack_from("<Q", raw, 32776)[0]
        assert abs(last_seen - (time.time() - 1) * 1e6) < 5e6
        idx, info, util = E.nvml_snapshot()
        uuid = next(u for u, i in idx.items() if i == 0)
        assert info[uuid].get(holder.pid, 0) >= 256 << 20
        cfg = E.parse_config(b"\0" * 1848)
        cfg["devices"][0].update({"activate": 1, "uuid_raw": uuid.encode().ljust(48, b"\0"), "total_memory": 4 << 30, "real_memory": 4 << 30})
        samples = {m: v for m, _, v in E.container_metrics(cfg, [holder.pid], idx, info, util, "node")}
        assert samples["container_vgpu_device_physical_memory_usage_in_bytes"] == info[uuid][holder.pid]
        assert samples["container_vgpu_device_memory_limit_in_bytes"] == float(4 << 30)
    finally:
        holder.wait(timeout=60)
