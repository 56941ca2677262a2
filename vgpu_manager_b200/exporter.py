"""Monitor-side consumer of the contract files.

The reference's device-monitor is Go (pkg/metrics/collector/node_gpu.go); this is the part of it
that reads what the interception library writes, restated so that the files produced by the B200
library can be shown to yield the same per-container metrics:

    container_vgpu_device_memory_limit_in_bytes              node_gpu.go:166-170, emitted :598-603
    container_vgpu_device_physical_memory_limit_in_bytes     :171-175, :604-609
    container_vgpu_device_memory_usage_in_bytes              :177-181, :635-640  (NVML usage + UVA ledger)
    container_vgpu_device_physical_memory_usage_in_bytes     :182-186, :641-646
    container_vgpu_device_memory_utilization_percent         :188-192, :648-653
    container_vgpu_device_core_utilization_percent           :193-197, :654-659

Inputs: the container's vgpu.config (1848 B resource_data_t), the node's vmem_node.config
(262 272 B, read under the per-device byte-range lock like vmem_config.go:179-205), NVML's
per-process figures and the container's pid set.  `python -m vgpu_manager_b200.exporter --help`
prints the Prometheus text exposition for one container using the real NVML.
"""
import argparse
import fcntl
import os
import struct
import sys

CFG_SIZE, VMEM_SIZE = 1848, 262272
MAX_DEVICES, MAX_PIDS = 16, 1024
DEV_OFF, DEV_STRIDE = 248, 96
VMEM_DEV_STRIDE, VMEM_SIZE_OFF, VMEM_LOCK_OFF = 16392, 16384, 16388
LABELS = ("pod_namespace", "pod_name", "container_name", "vdevice_idx", "device_uuid", "node")


def _cstr(b):
    return b.split(b"\0", 1)[0].decode("utf-8", "replace")


def parse_config(raw):
    """resource_data_t (hook.h:161-189) -> dict; raises on a wrong size like the Go mmap helper."""
    if len(raw) != CFG_SIZE:
        raise ValueError("vgpu.config must be %d bytes, got %d" % (CFG_SIZE, len(raw)))
    cfg = {"pod_uid": _cstr(raw[8:56]), "pod_name": _cstr(raw[56:120]), "pod_namespace": _cstr(raw[120:184]),
           "container_name": _cstr(raw[184:248]), "devices": []}
    for i in range(MAX_DEVICES):
        o = DEV_OFF + i * DEV_STRIDE
        uuid = raw[o:o + 48]
        total, real = struct.unpack_from("<QQ", raw, o + 48)
        hard, soft, core_limit, hard_limit, mem_limit, oversold, activate = struct.unpack_from("<7i", raw, o + 64)
        cfg["devices"].append({"uuid_raw": uuid, "total_memory": total, "real_memory": real, "hard_core": hard,
                               "soft_core": soft, "activate": activate})
    cfg["compatibility_mode"], cfg["sm_watcher"], cfg["vmem_node"] = struct.unpack_from("<3i", raw, 1784)
    return cfg


def ledger_usage(path, host_index):
    """Sum of `used` over the device's records, under F_RDLCK on its lock byte (node_gpu.go:612-632)."""
    if host_index < 0 or host_index >= MAX_DEVICES:
        return 0
    try:
        fd = os.open(path, os.O_RDONLY)
    except OSError:
        return 0
    try:
        if os.fstat(fd).st_size != VMEM_SIZE:
            return 0
        base = host_index * VMEM_DEV_STRIDE
        lock = struct.pack("hhqqi", fcntl.F_RDLCK, os.SEEK_SET, base + VMEM_LOCK_OFF, 1, 0)
        fcntl.fcntl(fd, fcntl.F_SETLKW, lock)
        try:
            raw = os.pread(fd, VMEM_DEV_STRIDE, base)
        finally:
            fcntl.fcntl(fd, fcntl.F_SETLK, struct.pack("hhqqi", fcntl.F_UNLCK, os.SEEK_SET, base + VMEM_LOCK_OFF, 1, 0))
    finally:
        os.close(fd)
    n = min(struct.unpack_from("<I", raw, VMEM_SIZE_OFF)[0], MAX_PIDS)
    return sum(struct.unpack_from("<iiQ", raw, 16 * i)[2] for i in range(n)) & 0xFFFFFFFFFFFFFFFF


def _valid(x):          # util.GetValidValue
    return x if x <= 100 else 0


def container_metrics(cfg, container_pids, dev_index_map, proc_info, proc_util, node, vmem_path=None, vmem_enabled=False):
    """The per-container block of Collect() (node_gpu.go:546-660).

    dev_index_map: {uuid: host index}; proc_info: {uuid: {pid: usedGpuMemory}};
    proc_util: {uuid: {pid: (sm, enc, dec)}}.  Returns [(metric, labels dict, value)] in emission order."""
    out = []
    count = 0
    for i in range(MAX_DEVICES):
        dev = cfg["devices"][i]
        if dev["activate"] == 0:
            continue
        try:
            uuid = dev["uuid_raw"][:40].decode("utf-8")
        except UnicodeDecodeError:
            continue
        if uuid not in dev_index_map:
            continue
        host_index = dev_index_map[uuid]
        limit, real = dev["total_memory"], dev["real_memory"]
        vidx = str(count)
        count += 1
        mem = vmem = sm_util = 0
        gpu_pids = []
        infos, utils = proc_info.get(uuid) or {}, proc_util.get(uuid) or {}
        for pid in container_pids:
            if pid in infos:
                gpu_pids.append(pid)
                mem = (mem + infos[pid]) & 0xFFFFFFFFFFFFFFFF
        for pid in container_pids:
            if pid in utils:
                sm, enc, dec = utils[pid]
                sm_util = (sm_util + _valid(sm) + (_valid(enc) + _valid(dec)) * 85 // 100) & 0xFFFFFFFF
        labels = dict(zip(LABELS, (cfg["pod_namespace"], cfg["pod_name"], cfg["container_name"], vidx, uuid, node)))
        out.append(("container_vgpu_device_memory_limit_in_bytes", labels, float(limit)))
        out.append(("container_vgpu_device_physical_memory_limit_in_bytes", labels, float(real)))
        if vmem_enabled and gpu_pids and vmem_path:
            vmem = ledger_usage(vmem_path, host_index)
        out.append(("container_vgpu_device_memory_usage_in_bytes", labels, float((mem + vmem) & 0xFFFFFFFFFFFFFFFF)))
        out.append(("container_vgpu_device_physical_memory_usage_in_bytes", labels, float(mem)))
        usage = (mem + vmem) & 0xFFFFFFFFFFFFFFFF
        if usage >= limit:
            rate = 100
        elif limit > 0:
            rate = int(float(usage) / float(limit) * 100)
        else:
            rate = 0
        out.append(("container_vgpu_device_memory_utilization_percent", labels, float(rate)))
        out.append(("container_vgpu_device_core_utilization_percent", labels, float(min(sm_util, 100))))
    return out


def exposition(samples):
    lines = []
    for name, labels, value in samples:
        lab = ",".join('%s="%s"' % (k, str(v).replace("\\", "\\\\").replace('"', '\\"')) for k, v in labels.items())
        lines.append("%s{%s} %s" % (name, lab, repr(value) if value != int(value) else "%d" % value))
    return "\n".join(lines) + ("\n" if lines else "")


def nvml_snapshot():
    """{uuid: host index}, per-process memory and utilisation maps from the real NVML."""
    import time

    import pynvml
    pynvml.nvmlInit()
    idx, info, util = {}, {}, {}
    for i in range(pynvml.nvmlDeviceGetCount()):
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        uuid = pynvml.nvmlDeviceGetUUID(h)
        uuid = uuid.decode() if isinstance(uuid, bytes) else uuid
        idx[uuid] = i
        procs = {}
        for getter in (pynvml.nvmlDeviceGetComputeRunningProcesses, pynvml.nvmlDeviceGetGraphicsRunningProcesses):
            try:
                for p in getter(h):
                    procs.setdefault(int(p.pid), int(p.usedGpuMemory or 0))
            except pynvml.NVMLError:
                pass
        info[uuid] = procs
        try:
            since = int((time.time() - 1.0) * 1e6)
            util[uuid] = {int(s.pid): (int(s.smUtil), int(s.encUtil), int(s.decUtil))
                          for s in pynvml.nvmlDeviceGetProcessUtilization(h, since)}
        except pynvml.NVMLError:
            util[uuid] = {}
    pynvml.nvmlShutdown()
    return idx, info, util


def serve(port, collect, host="0.0.0.0"):
    """Minimal /metrics endpoint (Prometheus text exposition); `collect()` returns the samples."""
    import http.server

    class Handler(http.server.BaseHTTPRequestHandler):
        def do_GET(self):
            if self.path.split("?")[0] not in ("/metrics", "/"):
                self.send_error(404)
                return
            try:
                body = exposition(collect()).encode()
            except Exception as e:  # a scrape must never kill the exporter
                self.send_error(500, str(e))
                return
            self.send_response(200)
            self.send_header("Content-Type", "text/plain; version=0.0.4; charset=utf-8")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *a):
            pass

    return http.server.ThreadingHTTPServer((host, port), Handler)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--config", default="/etc/vgpu-manager/config/vgpu.config")
    ap.add_argument("--vmem", default="/tmp/.vmem_node/vmem_node.config")
    ap.add_argument("--pids", required=True, help="comma-separated host pids of the container (or @file, one per line)")
    ap.add_argument("--node", default=os.uname().nodename)
    ap.add_argument("--no-vmem", action="store_true", help="VMemoryNode feature gate off")
    ap.add_argument("--listen", type=int, default=0, help="serve /metrics on this port instead of printing once")
    a = ap.parse_args(argv)
    if a.pids.startswith("@"):
        with open(a.pids[1:]) as f:
            pids = [int(x) for x in f.read().split()]
    else:
        pids = [int(x) for x in a.pids.split(",") if x]
    def collect():
        with open(a.config, "rb") as f:
            cfg = parse_config(f.read())
        idx, info, util = nvml_snapshot()
        return container_metrics(cfg, pids, idx, info, util, a.node, a.vmem, vmem_enabled=not a.no_vmem and bool(cfg["vmem_node"]))

    if a.listen:
        serve(a.listen, collect).serve_forever()
    sys.stdout.write(exposition(collect()))


if __name__ == "__main__":
    main()
