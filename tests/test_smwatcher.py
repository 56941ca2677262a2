"""vgpu-smwatcher (csrc/smwatcher.c), the node-level producer of sm_util.config.

Its oracle is a Python restatement of the Go producer (reference pkg/device/manager/watcher.go:128-184
smWatcherSingleDevice, file layout pkg/config/watcher/sm_watcher.go:34-63) applied to the same
scripted NVML; the consumers are the two interception libraries in external-watcher mode
(library/src/cuda_hook.c:1009-1042)."""
import ctypes as C
import fcntl
import os
import struct
import subprocess
import time

import pytest

import helpers as H

WATCHER = os.path.join(H.ROOT, "vgpu_manager_b200", "vgpu-smwatcher")
NVML_STUB = os.path.join(H.STUB_DIR, "libnvidia-ml.so.1")
FILE_SIZE = 1311232
DEV_SIZE = 81952
OTHERS = "111:1048576:c:30,222:2097152:cg:10,333:4096:g,444:8192:c"


class ProcV2(C.Structure):
    _fields_ = [("pid", C.c_uint32), ("_pad", C.c_uint32), ("used", C.c_uint64), ("gi", C.c_uint32), ("ci", C.c_uint32)]


class SmDev(C.Structure):
    _fields_ = [("samples", H.UtilSample * 1024), ("samples_size", C.c_uint32), ("_pad0", C.c_uint32),
                ("last_seen_us", C.c_uint64), ("compute", ProcV2 * 1024), ("compute_size", C.c_uint32),
                ("_pad1", C.c_uint32), ("graphics", ProcV2 * 1024), ("graphics_size", C.c_uint32),
                ("lock_byte", C.c_uint8), ("_pad2", C.c_uint8 * 3)]


assert C.sizeof(SmDev) == DEV_SIZE and SmDev.lock_byte.offset == 81948


def expected_device(others):
    """smWatcherSingleDevice on the scripted NVML: lists in NVML order, samples for compute pids with sm > 0."""
    comp, graph, samples = [], [], []
    for tok in others.split(","):
        f = tok.split(":")
        pid, used, kind = int(f[0]), int(f[1]), f[2]
        sm = int(f[3]) if len(f) > 3 else 0
        if "c" in kind:
            comp.append((pid, used))
            if sm:
                samples.append((pid, sm))
        if "g" in kind:
            graph.append((pid, used))
    return comp, graph, samples


def run_watcher(path, passes, env_extra=None, period_ms=10, extra=()):
    env = dict(os.environ, STUB_GPU_COUNT="2", STUB_OTHER_PROCS=OTHERS)
    env.update(env_extra or {})
    return subprocess.run([WATCHER, "--file", path, "--passes", str(passes), "--period-ms", str(period_ms), "--nvml", NVML_STUB, *extra],
                          env=env, capture_output=True, text=True, timeout=60)


@pytest.fixture(scope="module")
def built():
    H.build_all()
    assert os.path.exists(WATCHER), "vgpu-smwatcher was not built"


def test_file_is_created_sized_and_filled_like_the_go_producer(built, tmp_path):
    path = str(tmp_path / "watcher" / "sm_util.config")
    t0 = int(time.time() * 1e6)
    r = run_watcher(path, 2)
    t1 = int(time.time() * 1e6)
    assert r.returncode == 0, r.stderr
    assert os.path.getsize(path) == FILE_SIZE and (os.stat(path).st_mode & 0o777) == 0o644
    raw = open(path, "rb").read()
    comp, graph, samples = expected_device(OTHERS)
    for i in range(2):
        d = SmDev.from_buffer_copy(raw[i * DEV_SIZE:(i + 1) * DEV_SIZE])
        assert d.compute_size == len(comp) and d.graphics_size == len(graph) and d.samples_size == len(samples)
        assert [(p.pid, p.used) for p in d.compute[:d.compute_size]] == comp
        assert [(p.pid, p.used) for p in d.graphics[:d.graphics_size]] == graph
        assert all(p.gi == 0xFFFFFFFF and p.ci == 0xFFFFFFFF for p in d.compute[:d.compute_size])  # not a MIG instance (the fake NVML serves the 24-byte v3 records)
        assert [(s.pid, s.sm) for s in d.samples[:d.samples_size]] == samples
        # lastSeenTimeStamp = now - 1 s at the time of the pass; sample stamps are newer than it
        assert t0 - 1_000_000 <= d.last_seen_us <= t1 - 1_000_000
        assert all(s.ts_us > d.last_seen_us for s in d.samples[:d.samples_size])
        assert d.lock_byte == 0
    # devices beyond the NVML count are never touched
    assert raw[2 * DEV_SIZE:] == b"\0" * (FILE_SIZE - 2 * DEV_SIZE)


def test_wrong_sized_file_is_recreated_and_stale_samples_survive_a_failed_query(built, tmp_path):
    path = str(tmp_path / "sm_util.config")
    with open(path, "wb") as f:
        f.write(b"\xff" * 1000)
    assert run_watcher(path, 1).returncode == 0
    assert os.path.getsize(path) == FILE_SIZE
    before = SmDev.from_buffer_copy(open(path, "rb").read()[:DEV_SIZE])
    assert before.samples_size == 2
    # no process has any utilisation -> NVML answers NOT_FOUND: lists are refreshed, the samples are kept (watcher.go:176-182)
    r = run_watcher(path, 1, {"STUB_OTHER_PROCS": "111:1048576:c,555:4096:c"})
    assert r.returncode == 0, r.stderr
    after = SmDev.from_buffer_copy(open(path, "rb").read()[:DEV_SIZE])
    assert [p.pid for p in after.compute[:after.compute_size]] == [111, 555]
    assert after.samples_size == 2 and [s.pid for s in after.samples[:2]] == [s.pid for s in before.samples[:2]]
    assert after.last_seen_us >= before.last_seen_us


def test_device_record_is_written_under_its_byte_range_lock(built, tmp_path):
    """A reader holding F_RDLCK on device 0's lock byte (what the library does, cuda_hook.c:1016) blocks
    the producer for exactly that device; device 1's byte is independent."""
    path = str(tmp_path / "sm_util.config")
    assert run_watcher(path, 1).returncode == 0
    fd = os.open(path, os.O_RDONLY)
    off0 = SmDev.lock_byte.offset
    fcntl.fcntl(fd, fcntl.F_SETLK, struct.pack("hhqqi", fcntl.F_RDLCK, os.SEEK_SET, off0, 1, 0))
    env = dict(os.environ, STUB_GPU_COUNT="2", STUB_OTHER_PROCS="777:4096:c:5")
    p = subprocess.Popen([WATCHER, "--file", path, "--passes", "1", "--period-ms", "10", "--nvml", NVML_STUB], env=env)
    time.sleep(0.5)
    assert p.poll() is None, "the producer did not wait for the reader's lock"
    d0 = SmDev.from_buffer_copy(open(path, "rb").read()[:DEV_SIZE])
    assert d0.compute[0].pid == 111  # untouched while we hold the lock
    fcntl.fcntl(fd, fcntl.F_SETLK, struct.pack("hhqqi", fcntl.F_UNLCK, os.SEEK_SET, off0, 1, 0))
    assert p.wait(timeout=10) == 0
    os.close(fd)
    raw = open(path, "rb").read()
    assert SmDev.from_buffer_copy(raw[:DEV_SIZE]).compute[0].pid == 777
    assert SmDev.from_buffer_copy(raw[DEV_SIZE:2 * DEV_SIZE]).compute[0].pid == 777


def test_both_libraries_consume_the_produced_file(built):
    """External-watcher mode end to end: the daemon publishes, a tenant under either library starts,
    reports the same numbers and runs a launch train with the limiter fed from the file."""
    base = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": H.STUB_UUID, "LOGGER_LEVEL": "1",
            "CUDA_MEM_LIMIT_0": "1g", "CUDA_CORE_LIMIT_0": "30", "EXTERNAL_SM_WATCHER_ENABLED": "true",
            "STUB_OTHER_PROCS": OTHERS, "STUB_UTIL": "fixed:40"}
    outs = []
    for lib in (H.REF_SO, H.NEW_SO):
        if not os.path.exists(lib):
            continue
        sb = H.Sandbox()
        path = sb.path("etc/vgpu-manager/watcher/sm_util.config")
        daemon = subprocess.Popen([WATCHER, "--file", path, "--period-ms", "20", "--nvml", NVML_STUB],
                                  env=dict(os.environ, STUB_GPU_COUNT="1", STUB_OTHER_PROCS=OTHERS))
        try:
            for _ in range(100):
                if os.path.exists(path) and os.path.getsize(path) == FILE_SIZE:
                    break
                time.sleep(0.02)
            out, err, rc = H.run_scenario(lib, "init 0\ntotalmem\nmeminfo\nlaunch 2000\nsleepms 300\nlaunch 2000\nmeminfo\n", base,
                                          sb=sb, check=False)
        finally:
            daemon.terminate()
            daemon.wait(timeout=10)
        outs.append((out, rc))
        sb.cleanup()
    assert all(rc == 0 for _, rc in outs), outs
    assert len(outs) < 2 or outs[0] == outs[1]


# ----------------------------------------------------------------------------------------------------------------
# samples from the tenants' on-device readings (VGPU_LOCK_DIR/vgpu_<i>.readings) instead of
# nvmlDeviceGetProcessUtilization


class Reading(C.Structure):
    _fields_ = [("owner", C.c_uint64), ("ts_us", C.c_uint64), ("sm_pct", C.c_uint32), ("queue_busy_pct", C.c_uint32),
                ("sm_active_pct", C.c_uint32), ("seq", C.c_uint32)]


assert C.sizeof(Reading) == 32
READINGS_SIZE = 32 * 1024


def owner_key(pid):
    """What a tenant publishes under: (inode of its pid namespace << 32) | its pid inside that namespace."""
    inner = pid
    for line in open("/proc/%d/status" % pid):
        if line.startswith("NSpid:"):
            inner = int(line.split()[-1])
    return (os.stat("/proc/%d/ns/pid" % pid).st_ino << 32) | inner


def write_readings(path, entries):
    raw = bytearray(READINGS_SIZE)
    for slot, (owner, ts, sm) in entries.items():
        struct.pack_into("<QQIIII", raw, 32 * slot, owner, ts, sm, sm, 0, 1)
    with open(path, "wb") as f:
        f.write(raw)


def read_slots(path):
    raw = open(path, "rb").read()
    return [r for r in (Reading.from_buffer_copy(raw[32 * i:32 * i + 32]) for i in range(len(raw) // 32)) if r.owner]


def test_samples_are_built_from_the_tenants_own_readings(built, tmp_path):
    """--source device: a compute process with a fresh reading gets a sample carrying that reading (not NVML's figure);
    processes that publish nothing, stale readings and readings of pids NVML does not list produce none.
    --source mixed: NVML's samples, with the tenant's own reading replacing NVML's sample of its pid."""
    me = os.getpid()
    sleeper = subprocess.Popen(["sleep", "30"])
    try:
        now = int(time.time() * 1e6)
        rd = tmp_path / "lock"
        rd.mkdir()
        write_readings(str(rd / "vgpu_0.readings"), {5: (owner_key(me), now, 42),                      # fresh, listed by NVML
                                                     9: (owner_key(sleeper.pid), now - 5_000_000, 77),  # stale
                                                     11: ((1 << 32) | 999999, now, 88)})                 # nobody NVML lists
        write_readings(str(rd / "vgpu_1.readings"), {0: (owner_key(sleeper.pid), now, 13)})
        procs = "%d:1048576:c:30,%d:2097152:c:10,333:4096:g" % (me, sleeper.pid)
        path = str(tmp_path / "sm_util.config")
        r = run_watcher(path, 1, {"STUB_OTHER_PROCS": procs}, extra=["--source", "device", "--readings-dir", str(rd)])
        assert r.returncode == 0, r.stderr
        raw = open(path, "rb").read()
        d0, d1 = (SmDev.from_buffer_copy(raw[i * DEV_SIZE:(i + 1) * DEV_SIZE]) for i in range(2))
        assert [(p.pid, p.used) for p in d0.compute[:d0.compute_size]] == [(me, 1048576), (sleeper.pid, 2097152)]  # lists: NVML
        assert [(s.pid, s.sm, s.mem, s.enc, s.dec) for s in d0.samples[:d0.samples_size]] == [(me, 42, 0, 0, 0)]
        assert d0.samples[0].ts_us == now and d0.samples[0].ts_us > d0.last_seen_us
        assert [(s.pid, s.sm) for s in d1.samples[:d1.samples_size]] == [(sleeper.pid, 13)]
        # nothing fresh any more (NVML's NOT_FOUND case): the lists are refreshed, the previous samples stay
        time.sleep(1.1)
        r = run_watcher(path, 1, {"STUB_OTHER_PROCS": procs}, extra=["--source", "device", "--readings-dir", str(rd)])
        assert r.returncode == 0, r.stderr
        again = SmDev.from_buffer_copy(open(path, "rb").read()[:DEV_SIZE])
        assert again.samples_size == 1 and again.samples[0].ts_us == now and again.last_seen_us > d0.last_seen_us
        # mixed
        now = int(time.time() * 1e6)
        write_readings(str(rd / "vgpu_0.readings"), {5: (owner_key(me), now, 42)})
        r = run_watcher(path, 1, {"STUB_OTHER_PROCS": procs}, extra=["--source", "mixed", "--readings-dir", str(rd)])
        assert r.returncode == 0, r.stderr
        m0 = SmDev.from_buffer_copy(open(path, "rb").read()[:DEV_SIZE])
        assert sorted((s.pid, s.sm) for s in m0.samples[:m0.samples_size]) == sorted([(me, 42), (sleeper.pid, 10)])
    finally:
        sleeper.kill()
        sleeper.wait()


def test_a_tenant_on_an_on_device_signal_publishes_its_reading_and_the_watcher_serves_it(built):
    """End to end on the fake driver: a tenant whose limiter steers on the stream queue-busy signal publishes one
    reading per control period next to the GPU lock file; vgpu-smwatcher --source device turns it into the tenant's
    sample of sm_util.config without asking NVML for utilisation (the fake NVML would have said 77).  A tenant on the
    default NVML reading publishes nothing."""
    base = {"MANAGER_COMPATIBILITY_MODE": "0", "MANAGER_VISIBLE_DEVICES": H.STUB_UUID, "LOGGER_LEVEL": "1",
            "CUDA_CORE_LIMIT_0": "30", "STUB_UTIL": "fixed:40"}
    script = "init 0\n" + "launch 500\nsleepms 100\n" * 80  # ~8 s; the tenant is killed as soon as the test has seen enough
    for source, expect in (("queue", True), (None, False)):
        sb = H.Sandbox()
        env = dict(base)
        if source:
            env["VGPU_B200_UTIL_SOURCE"] = source
        tenant = subprocess.Popen([H.SCENARIO], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                  env=H.preload_env(H.NEW_SO, sb, env))
        try:
            tenant.stdin.write(script)
            tenant.stdin.close()
            rfile = sb.path("lock/vgpu_0.readings")
            t_end = time.time() + (7.0 if expect else 1.5)  # generous on a loaded machine; the positive case leaves early
            slots = []
            while time.time() < t_end and not (slots and slots[0].seq >= 3):
                time.sleep(0.1)
                slots = read_slots(rfile) if os.path.exists(rfile) else []
            if not expect:
                assert not slots, [(s.owner, s.seq) for s in slots]
                continue
            assert len(slots) == 1 and slots[0].owner == owner_key(tenant.pid), [(hex(s.owner), s.seq) for s in slots]
            assert slots[0].seq >= 3 and slots[0].sm_pct <= 100 and abs(slots[0].ts_us - time.time() * 1e6) < 2e6
            path = sb.path("etc/vgpu-manager/watcher/sm_util.config")
            r = subprocess.run([WATCHER, "--file", path, "--passes", "1", "--nvml", NVML_STUB, "--source", "device",
                                "--readings-dir", sb.path("lock")],
                               env=dict(os.environ, STUB_GPU_COUNT="1", STUB_OTHER_PROCS="%d:1048576:c:77" % tenant.pid),
                               capture_output=True, text=True, timeout=30)
            assert r.returncode == 0, r.stderr
            d = SmDev.from_buffer_copy(open(path, "rb").read()[:DEV_SIZE])
            fresh = read_slots(rfile)[0]
            assert d.samples_size == 1 and d.samples[0].pid == tenant.pid and d.samples[0].sm <= 100 and d.samples[0].sm != 77
            assert abs(int(d.samples[0].sm) - int(fresh.sm_pct)) <= 100 and d.samples[0].ts_us >= d.last_seen_us
        finally:
            tenant.kill()
            tenant.wait()
            sb.cleanup()
