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


def test_on_device_readings_reach_sm_util_config(built, tmp_path):
    """SURVEY.md 8f-1 on the real driver: a 50 %-capped busy tenant whose limiter steers on the stream queue-busy
    signal publishes its reading once per control period next to the GPU lock file; vgpu-smwatcher --source device
    serves it as the tenant's sample in sm_util.config without calling nvmlDeviceGetProcessUtilization.  NVML's own
    per-process figure of the same tenant (a second pass with --source nvml) is recorded beside it for profiles/."""
    import json

    import band
    import test_smwatcher as T
    sb = H.Sandbox()
    env = band.tenant_env(H.NEW_SO, sb, 50)
    env["VGPU_B200_UTIL_SOURCE"] = "queue"
    tenant = subprocess.Popen([H.STORM, "--steps", "1000000", "--warmup", "0", "--per-step", "200", "--max-seconds", "12", *band.BUSY],
                              env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        rfile = sb.path("lock/vgpu_0.readings")
        slots, t_end = [], time.time() + 8
        while time.time() < t_end and not (slots and slots[0].seq >= 12):  # ~1 s of control periods after bring-up
            time.sleep(0.2)
            slots = T.read_slots(rfile) if os.path.exists(rfile) else []
        assert len(slots) == 1 and slots[0].owner == T.owner_key(tenant.pid) and slots[0].seq >= 12, [(hex(s.owner), s.seq) for s in slots]
        path = str(tmp_path / "sm_util.config")
        r = subprocess.run([WATCHER, "--file", path, "--passes", "2", "--period-ms", "100", "--source", "device", "--readings-dir",
                            sb.path("lock")], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        d = T.SmDev.from_buffer_copy(open(path, "rb").read()[:T.DEV_SIZE])
        mine = [s for s in d.samples[:d.samples_size] if s.pid == tenant.pid]
        assert len(mine) == 1 and mine[0].sm <= 100 and mine[0].ts_us >= d.last_seen_us, [(s.pid, s.sm) for s in d.samples[:d.samples_size]]
        assert tenant.pid in [p.pid for p in d.compute[:d.compute_size]]
        path2 = str(tmp_path / "sm_util_nvml.config")
        r2 = subprocess.run([WATCHER, "--file", path2, "--passes", "2", "--period-ms", "100"], capture_output=True, text=True, timeout=60)
        nv = []
        if r2.returncode == 0:
            d2 = T.SmDev.from_buffer_copy(open(path2, "rb").read()[:T.DEV_SIZE])
            nv = [int(s.sm) for s in d2.samples[:d2.samples_size] if s.pid == tenant.pid]
        os.makedirs(os.path.join(H.ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(H.ROOT, "gpurun_out", "smwatcher_device_readings.json"), "w") as f:
            json.dump({"tenant_pid": tenant.pid, "cap_pct": 50, "device_reading_sm_pct": int(mine[0].sm), "nvml_sample_sm_pct": nv,
                       "readings_published": int(T.read_slots(rfile)[0].seq)}, f)
    finally:
        tenant.wait(timeout=120)
        sb.cleanup()
