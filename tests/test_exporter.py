"""vgpu_manager_b200/exporter.py - the monitor-side consumer (reference
pkg/metrics/collector/node_gpu.go:546-660) applied to the files each library leaves behind: the
B200 library's vgpu.config + vmem_node.config must yield the same per-container metrics as the
reference library's."""
import os
import struct

import pytest

import helpers as H
from vgpu_manager_b200 import exporter as E

MiB, GiB = 1 << 20, 1 << 30
BASE = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": H.STUB_UUID, "LOGGER_LEVEL": "0",
        "VGPU_POD_NAME": "trainer-0", "VGPU_POD_NAMESPACE": "ml", "VGPU_POD_UID": "8d0a2c1e-5b7f-4c55-9d3e-0f2b6f1a7c11",
        "VGPU_CONTAINER_NAME": "main"}


@pytest.fixture(scope="module")
def built():
    H.build_all()


def ledger_pids(path, host_index=0):
    raw = open(path, "rb").read()
    base = host_index * 16392
    n = struct.unpack_from("<I", raw, base + 16384)[0]
    return [struct.unpack_from("<iiQ", raw, base + 16 * i)[0] for i in range(n)]


def metrics_after(lib, script, env):
    """Metrics computed while the tenant is still alive (both libraries drop their ledger record at exit)."""
    import subprocess
    import time
    sb = H.Sandbox()
    e = dict(BASE)
    e.update(env)
    p = subprocess.Popen([H.SCENARIO], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         env=H.preload_env(lib, sb, e))
    p.stdin.write(script + "ledger 0\nsleepms 1500\n")
    p.stdin.flush()
    line = ""
    while not line.startswith("ledger"):  # the tenant has issued every allocation once this line arrives
        line = p.stdout.readline()
        assert line, p.stderr.read()[-1500:]
    try:
        return _snapshot(sb)
    finally:
        p.stdin.close()
        assert p.wait(timeout=30) == 0
        sb.cleanup()


def _snapshot(sb):
    cfg = E.parse_config(sb.config_bytes())
    pids = ledger_pids(sb.ledger()) if os.path.exists(sb.ledger()) else []
    assert len(pids) <= 1
    pid = pids[0] if pids else 4242
    samples = E.container_metrics(cfg, [pid, 999999], {H.STUB_UUID: 0}, {H.STUB_UUID: {pid: 700 * MiB, 31337: GiB}},
                                  {H.STUB_UUID: {pid: (37, 200, 10), 31337: (50, 0, 0)}}, "node-a", sb.ledger(),
                                  vmem_enabled=bool(cfg["vmem_node"]))
    return samples, pid


def test_same_metrics_from_either_librarys_files(built):
    script = "init 0\n" + "alloc %d\n" % (64 * MiB) * 40 + "free 35\nmanaged %d 1\n" % (3 * MiB)
    env = {"CUDA_MEM_LIMIT_0": "8g", "CUDA_MEM_RATIO_0": "4", "VMEMORY_NODE_ENABLED": "true"}
    got = {}
    for name, lib in (("reference", H.REF_SO), ("b200", H.NEW_SO)):
        if os.path.exists(lib):
            got[name] = metrics_after(lib, script, env)[0]
    b = got["b200"]
    if "reference" in got:
        assert got["reference"] == b
    by = {m: v for m, _, v in b}
    labels = b[0][1]
    assert labels == {"pod_namespace": "ml", "pod_name": "trainer-0", "container_name": "main", "vdevice_idx": "0",
                      "device_uuid": H.STUB_UUID, "node": "node-a"}
    assert by["container_vgpu_device_memory_limit_in_bytes"] == 8 * GiB
    assert by["container_vgpu_device_physical_memory_limit_in_bytes"] == 2 * GiB
    assert by["container_vgpu_device_physical_memory_usage_in_bytes"] == 700 * MiB
    # 32 allocations fit the 2 GiB of physical memory, 8 spilled to UVA, one of those freed, + 3 MiB managed
    spilled = 7 * 64 * MiB + 3 * MiB
    assert by["container_vgpu_device_memory_usage_in_bytes"] == 700 * MiB + spilled
    assert by["container_vgpu_device_memory_utilization_percent"] == int((700 * MiB + spilled) / (8 * GiB) * 100)
    assert by["container_vgpu_device_core_utilization_percent"] == 37 + (0 + 10) * 85 // 100  # enc 200 is invalid -> 0


def test_ledger_is_ignored_without_the_feature_gate_and_without_gpu_pids(built):
    script = "init 0\n" + "alloc %d\n" % (512 * MiB) * 6
    samples, _ = metrics_after(H.NEW_SO, script, {"CUDA_MEM_LIMIT_0": "4g", "CUDA_MEM_RATIO_0": "2"})
    by = {m: v for m, _, v in samples}
    assert by["container_vgpu_device_memory_usage_in_bytes"] == by["container_vgpu_device_physical_memory_usage_in_bytes"]
    cfg = E.parse_config(b"\0" * 1848)
    assert E.container_metrics(cfg, [1], {}, {}, {}, "n") == []
    with pytest.raises(ValueError):
        E.parse_config(b"\0" * 100)


def test_exposition_format():
    text = E.exposition([("m_bytes", {"a": 'x"y', "b": "1"}, 1024.0), ("m_ratio", {"a": "z"}, 0.5)])
    assert text == 'm_bytes{a="x\\"y",b="1"} 1024\nm_ratio{a="z"} 0.5\n'


def test_metrics_endpoint_serves_the_exposition():
    import threading
    import urllib.request
    samples = [("container_vgpu_device_memory_usage_in_bytes", {"pod_name": "p", "vdevice_idx": "0"}, 4096.0)]
    srv = E.serve(0, lambda: samples, host="127.0.0.1")
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    try:
        port = srv.server_address[1]
        with urllib.request.urlopen("http://127.0.0.1:%d/metrics" % port, timeout=5) as r:
            assert r.status == 200 and r.headers["Content-Type"].startswith("text/plain")
            assert r.read().decode() == 'container_vgpu_device_memory_usage_in_bytes{pod_name="p",vdevice_idx="0"} 4096\n'
        try:
            urllib.request.urlopen("http://127.0.0.1:%d/other" % port, timeout=5)
            assert False, "404 expected"
        except urllib.error.HTTPError as e:
            assert e.code == 404
    finally:
        srv.shutdown()
