"""Shared test helpers: builds, ctypes bindings of the oracle, scenario runner."""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.join(ROOT, "tests")
BUILD = os.path.join(TESTS, "_build")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvgpu-control.so")
REF_COSIM = os.path.join(ROOT, "oracle", "_ref", "ref_cosim")
NEW_SO = os.path.join(ROOT, "vgpu_manager_b200", "libvgpu-control.so")
STUB_DIR = os.path.join(BUILD, "stub")
REDIRECT = os.path.join(BUILD, "libredirect.so")
SCENARIO = os.path.join(BUILD, "scenario")
STORM = os.path.join(BUILD, "storm")
STUB_UUID = "GPU-11111111-1111-1111-1111-111111111111"

_built = False


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-4000:], r.stderr[-4000:]))


def build_all():
    global _built
    if _built:
        return
    _run(["make", "-s"], os.path.join(ROOT, "oracle"))
    _run(["make", "-s"], TESTS)
    import importlib.util
    spec = importlib.util.spec_from_file_location("vgpu_build", os.path.join(ROOT, "vgpu_manager_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    _built = True


def have_reference():
    return os.path.exists(REF_SO) and os.path.exists(REF_COSIM)


# ----------------------------------------------------------------------------- contract structs
class CfgDev(C.Structure):
    _fields_ = [("uuid", C.c_char * 48), ("total_memory", C.c_uint64), ("real_memory", C.c_uint64),
                ("hard_core", C.c_int32), ("soft_core", C.c_int32), ("core_limit", C.c_int32),
                ("hard_limit", C.c_int32), ("memory_limit", C.c_int32), ("memory_oversold", C.c_int32),
                ("activate", C.c_int32), ("_pad", C.c_int32)]


class Cfg(C.Structure):
    _fields_ = [("driver_major", C.c_int32), ("driver_minor", C.c_int32), ("pod_uid", C.c_char * 48),
                ("pod_name", C.c_char * 64), ("pod_namespace", C.c_char * 64), ("container_name", C.c_char * 64),
                ("devices", CfgDev * 16), ("compatibility_mode", C.c_int32), ("sm_watcher", C.c_int32),
                ("vmem_node", C.c_int32), ("reg_uuid", C.c_char * 48), ("_pad", C.c_int32)]


class Proc(C.Structure):
    _fields_ = [("pid", C.c_uint32), ("_pad", C.c_uint32), ("used_bytes", C.c_uint64)]


class UtilSample(C.Structure):
    _fields_ = [("pid", C.c_uint32), ("_pad", C.c_uint32), ("ts_us", C.c_uint64), ("sm", C.c_uint32),
                ("mem", C.c_uint32), ("enc", C.c_uint32), ("dec", C.c_uint32)]


class VmemRec(C.Structure):
    _fields_ = [("pid", C.c_int32), ("_pad", C.c_int32), ("used", C.c_uint64)]


class VmemDev(C.Structure):
    _fields_ = [("processes", VmemRec * 1024), ("processes_size", C.c_uint32), ("lock_byte", C.c_uint8),
                ("_pad", C.c_uint8 * 3)]


class OrcGpu(C.Structure):
    _fields_ = [("sm_num", C.c_int32), ("max_thread_per_sm", C.c_int32), ("total_cores", C.c_int64)]


class OrcUtil(C.Structure):
    _fields_ = [("user_current", C.c_int32), ("sys_current", C.c_int32), ("valid", C.c_int32),
                ("sys_process_num", C.c_int32)]


class OrcWatcher(C.Structure):
    _fields_ = [("share", C.c_int64), ("sys_free", C.c_int32), ("avg_sys_free", C.c_int32), ("i", C.c_int32),
                ("pre_sys_process_num", C.c_int32), ("up_limit", C.c_int32), ("_pad", C.c_int32)]


# kernel_abi.h mirrors
class QuotaReq(C.Structure):
    _fields_ = [("seq", C.c_uint32), ("kind", C.c_uint32), ("mode", C.c_uint32), ("allow_uva", C.c_uint32),
                ("memory_oversold", C.c_uint32), ("real_ok", C.c_uint32), ("n_compute", C.c_uint32),
                ("n_graphics", C.c_uint32), ("n_vmem", C.c_uint32), ("_pad", C.c_uint32),
                ("total_memory", C.c_uint64), ("real_memory", C.c_uint64), ("request", C.c_uint64),
                ("real_total", C.c_uint64), ("self_bytes", C.c_uint64), ("self_pid", C.c_uint32),
                ("_pad2", C.c_uint32 * 3), ("compute", Proc * 1024), ("graphics", Proc * 1024),
                ("vmem", VmemRec * 1024), ("cflags", C.c_uint8 * 1024), ("gflags", C.c_uint8 * 1024)]


class UtilReq(C.Structure):
    _fields_ = [("seq", C.c_uint32), ("status", C.c_uint32), ("mode", C.c_uint32), ("n_samples", C.c_uint32),
                ("sys_process_num", C.c_int32), ("have_container_pids", C.c_uint32), ("_pad", C.c_uint32 * 2),
                ("checktime_us", C.c_uint64), ("_pad2", C.c_uint64), ("samples", UtilSample * 1024),
                ("flags", C.c_uint8 * 1024)]


UTIL_NOTHING, UTIL_NPROC_ONLY, UTIL_SAMPLES = 0, 1, 2


def util_req_from_golden_step(mode, st, seq):
    """The publication the tick thread would have made for one step of a golden watcher trajectory
    (tests/golden/watcher.json: samples = [pid, sm, enc, dec, age_ms, mine]); the reference's NVML
    stand-in answers NOT_FOUND when there are no samples (oracle/ref_cosim.c)."""
    u = UtilReq()
    T = 1_700_000_000_000_000
    u.seq, u.mode, u.sys_process_num, u.have_container_pids = seq, mode, st["nproc"], 1
    u.checktime_us = T
    if not st["samples"]:
        u.status = UTIL_NPROC_ONLY
        return u
    u.status = UTIL_SAMPLES
    u.n_samples = len(st["samples"])
    for i, (pid, sm, enc, dec, age, mine) in enumerate(st["samples"]):
        s = u.samples[i]
        s.pid, s.sm, s.enc, s.dec = pid, sm, enc, dec
        s.ts_us = T + 1_000_000 - age * 1000
        u.flags[i] = 1 if mine else 0  # VGPU_FLAG_PRIMARY
    return u


class VslabReq(C.Structure):
    _fields_ = [("op", C.c_uint32), ("mask", C.c_uint32), ("want", C.c_uint32), ("set_mask", C.c_uint32), ("set_val", C.c_uint32),
                ("age", C.c_uint32), ("dptr", C.c_uint64), ("bytes", C.c_uint64), ("size", C.c_uint64), ("state", C.c_uint32),
                ("_pad", C.c_uint32)]


class VslabRes(C.Structure):
    _fields_ = [("dptr", C.c_uint64), ("bytes", C.c_uint64), ("size", C.c_uint64), ("state", C.c_uint32), ("age", C.c_uint32),
                ("slot", C.c_uint32), ("seq_done", C.c_uint32)]


VS_PUT, VS_TAKE, VS_SCAN, VS_UVA, VS_DEV = 0, 1, 2, 1, 2


def vslab_model_check(lib, seed=0x5EED, ops=400):
    """Random PUT / TAKE / SCAN(+flip) against a dictionary model of the placement table; returns the number of
    operations compared.  Used on the fake driver (plumbing) and on the GPU (the kernel itself)."""
    import random
    rng = random.Random(seed)
    model, age = {}, 0   # dptr -> [bytes, size, state, age]
    sizes = [2 << 20, 64 << 20, 256 << 20]
    n = 0
    for _ in range(ops):
        kind = rng.random()
        rq, rs = VslabReq(), VslabRes()
        if kind < 0.45 or not model:
            age += 1
            d = (rng.randrange(1, 1 << 20) << 21)
            size = rng.choice(sizes)
            st = rng.choice([VS_DEV, VS_UVA | VS_DEV, VS_UVA, 0])
            rq.op, rq.dptr, rq.bytes, rq.size, rq.state, rq.age = VS_PUT, d, size - rng.randrange(0, 4096), size, st, age
            lib.vslab_op(rq, rs)
            assert rs.slot != 0xFFFFFFFF
            model[d] = [rq.bytes, size, st, age]
        elif kind < 0.7:
            d = rng.choice(list(model)) if rng.random() < 0.8 else 0x123456000
            rq.op, rq.dptr = VS_TAKE, d
            lib.vslab_op(rq, rs)
            if d in model:
                assert (rs.dptr, rs.bytes, rs.size, rs.state) == (d, model[d][0], model[d][1], model[d][2]), (d, model[d])
                del model[d]
            else:
                assert rs.slot == 0xFFFFFFFF
        else:
            size = rng.choice(sizes)
            want = rng.choice([VS_DEV, VS_UVA | VS_DEV, 0])
            flip = VS_DEV if not (want & VS_DEV) else 0
            rq.op, rq.mask, rq.want, rq.set_mask, rq.set_val, rq.size = VS_SCAN, VS_UVA | VS_DEV, want, VS_DEV, flip, size
            lib.vslab_op(rq, rs)
            cands = [(v[3], d) for d, v in model.items() if v[1] == size and (v[2] & 3) == want]
            if not cands:
                assert rs.slot == 0xFFFFFFFF, (size, want, rs.dptr)
            else:
                _, d = min(cands)   # coldest = smallest age
                assert rs.dptr == d and rs.state == model[d][2], (rs.dptr, d)
                model[d][2] = (model[d][2] & ~VS_DEV) | flip
        n += 1
    # drain: everything that is left must come back exactly once
    for d in list(model):
        rq, rs = VslabReq(), VslabRes()
        rq.op, rq.dptr = VS_TAKE, d
        lib.vslab_op(rq, rs)
        assert rs.dptr == d and rs.state == model[d][2]
    return n


class QuotaRes(C.Structure):
    _fields_ = [("used", C.c_uint64), ("vmem", C.c_uint64), ("total", C.c_uint64), ("out_used", C.c_uint64),
                ("out_free", C.c_uint64), ("path", C.c_uint32), ("seq_done", C.c_uint32)]


class LimiterState(C.Structure):
    _fields_ = [("granted", C.c_longlong), ("consumed", C.c_longlong), ("bucket", C.c_longlong),
                ("share", C.c_longlong), ("up_limit", C.c_int), ("sys_free", C.c_int), ("avg_sys_free", C.c_int),
                ("ctr_i", C.c_int), ("pre_sys_process_num", C.c_int), ("valid", C.c_int),
                ("user_current", C.c_int), ("sys_current", C.c_int), ("sm_active_pct", C.c_int),
                ("queue_busy_pct", C.c_int), ("steps", C.c_ulonglong)]


_oracle = None


def oracle():
    """ctypes handle on oracle/liboracle.so (the checker)."""
    global _oracle
    if _oracle is None:
        build_all()
        o = C.CDLL(ORACLE_SO)
        o.orc_gpu_init.argtypes = [C.POINTER(OrcGpu), C.c_int, C.c_int]
        o.orc_delta.restype = C.c_int64
        o.orc_delta.argtypes = [C.POINTER(OrcGpu), C.c_int, C.c_int, C.c_int64]
        o.orc_change_token.restype = C.c_int64
        o.orc_change_token.argtypes = [C.POINTER(OrcGpu), C.c_int64, C.c_int64]
        o.orc_rate_limiter_try.argtypes = [C.POINTER(C.c_int64), C.c_uint32, C.c_uint32, C.c_uint32]
        o.orc_watcher_init.argtypes = [C.POINTER(OrcWatcher), C.POINTER(CfgDev)]
        o.orc_watcher_step.argtypes = [C.POINTER(OrcGpu), C.POINTER(CfgDev), C.POINTER(OrcWatcher),
                                       C.POINTER(OrcUtil), C.POINTER(C.c_int64)]
        o.orc_fold_utilization.argtypes = [C.c_int, C.POINTER(UtilSample), C.c_uint32, C.c_uint64,
                                           C.POINTER(C.c_uint8), C.POINTER(C.c_uint8), C.c_int, C.POINTER(OrcUtil)]
        o.orc_used_memory.restype = C.c_uint64
        o.orc_used_memory.argtypes = [C.c_int, C.POINTER(Proc), C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8),
                                      C.POINTER(Proc), C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
        o.orc_memory_path.argtypes = [C.POINTER(CfgDev), C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        o.orc_nvml_meminfo.argtypes = [C.POINTER(CfgDev), C.c_uint64, C.c_uint64] + [C.POINTER(C.c_uint64)] * 3
        o.orc_cu_meminfo.argtypes = [C.POINTER(CfgDev), C.c_uint64, C.c_uint64, C.c_int, C.c_uint64,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        o.orc_array_request.restype = C.c_uint64
        o.orc_array_request.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
        o.orc_array3d_request.restype = C.c_uint64
        o.orc_array3d_request.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
        o.orc_pitch_guess.restype = C.c_uint64
        o.orc_pitch_guess.argtypes = [C.c_uint64, C.c_uint32]
        o.orc_ledger_add.argtypes = [C.POINTER(VmemDev), C.c_int, C.c_uint64]
        o.orc_ledger_sub.argtypes = [C.POINTER(VmemDev), C.c_int, C.c_uint64]
        o.orc_ledger_sum.restype = C.c_uint64
        o.orc_ledger_sum.argtypes = [C.POINTER(VmemDev)]
        o.orc_ledger_rm_pid.argtypes = [C.POINTER(VmemDev), C.c_int]
        o.orc_ledger_purge.argtypes = [C.POINTER(VmemDev), C.c_int, C.POINTER(C.c_uint8)]
        o.orc_iec_to_bytes.restype = C.c_uint64
        o.orc_iec_to_bytes.argtypes = [C.c_char_p]
        o.orc_balance_batches.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _oracle = o
    return _oracle


GETENV_FN = C.CFUNCTYPE(C.c_char_p, C.c_char_p, C.c_void_p)


def oracle_config_from_env(env):
    """orc_config_from_env over a python dict -> 1848 raw bytes."""
    o = oracle()
    keep = {}

    def ge(name, ctx):
        k = name.decode()
        if k not in env:
            return None
        keep[k] = C.create_string_buffer(env[k].encode())
        return C.addressof(keep[k])

    fn = C.CFUNCTYPE(C.c_void_p, C.c_char_p, C.c_void_p)(ge)
    o.orc_config_from_env.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Cfg)]
    cfg = Cfg()
    o.orc_config_from_env(C.cast(fn, C.c_void_p), None, C.byref(cfg))
    return bytes(cfg)


def gpu_local_cpus(index=0):
    """CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function), or None.
    Tenants of both libraries are started there: a launch path that crosses the socket interconnect
    costs ~0.8 us per cuLaunchKernel on this class of box and would otherwise hit either arm at random."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        dom, rest = out.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/local_cpulist" % (dom[-4:].lower(), rest.lower())
        with open(path) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


def pin_to(cpus):
    """preexec_fn for subprocess: run the child on `cpus` (no-op for None)."""
    if not cpus:
        return None
    return lambda: os.sched_setaffinity(0, cpus)


# ----------------------------------------------------------------------------- process runners
class Sandbox:
    """Private copy of the contract directories, reached through the redirect shim."""

    def __init__(self):
        self.dir = tempfile.mkdtemp(prefix="vgpu_sb_", dir=os.path.join(BUILD))
        for d in ("etc/vgpu-manager/config", "etc/vgpu-manager/watcher", "etc/vgpu-manager/.host_proc", "lock", "vmem", "none"):
            os.makedirs(os.path.join(self.dir, d), exist_ok=True)
        self.extra_rules = []

    def rules(self, stub=True):
        r = ["/etc/vgpu-manager=%s/etc/vgpu-manager" % self.dir, "/tmp/.vgpu_lock=%s/lock" % self.dir,
             "/tmp/.vmem_node=%s/vmem" % self.dir]
        if stub:  # hide a real driver's version file so both libraries dlopen libcuda.so.1 (the stub)
            r.append("/proc/driver/nvidia=%s/none" % self.dir)
        return ":".join(r + self.extra_rules)

    def path(self, rel):
        return os.path.join(self.dir, rel)

    def ledger(self):
        return os.path.join(self.dir, "vmem", "vmem_node.config")

    def config_bytes(self):
        with open(self.path("etc/vgpu-manager/config/vgpu.config"), "rb") as f:
            return f.read()

    def cleanup(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def preload_env(lib, sb, env=None, stub=True):
    e = {k: v for k, v in os.environ.items() if not k.startswith(("CUDA_", "MANAGER_", "VGPU_", "STUB_", "LD_PRELOAD"))}
    e["VGPU_REDIRECT"] = sb.rules(stub)
    e["SCENARIO_LEDGER"] = sb.ledger()
    if stub:
        e["LD_LIBRARY_PATH"] = STUB_DIR
        # the fake GPU completes tenant kernels instantly, so "is tenant work executing" is never true there and
        # its utilisation is scripted instead (STUB_UTIL): keep every sampler window unless a test asks otherwise
        e["VGPU_B200_SKIP_IDLE_WINDOWS"] = "0"
    e["LD_PRELOAD"] = REDIRECT + ((" " + lib) if lib else "")
    e.update(env or {})
    return e


def run_scenario(lib, script, env=None, sb=None, stub=True, args=(), timeout=120, check=True):
    """Run tests/_build/scenario under `lib` (a preload .so or None). Returns (stdout, stderr, sandbox)."""
    build_all()
    own = sb is None
    sb = sb or Sandbox()
    r = subprocess.run([SCENARIO, *args], input=script, capture_output=True, text=True,
                       env=preload_env(lib, sb, env, stub), timeout=timeout)
    if r.returncode != 0 and check:
        raise RuntimeError("scenario failed rc=%d\n%s\n%s" % (r.returncode, r.stdout[-2000:], r.stderr[-2000:]))
    if not check:
        return r.stdout, r.stderr, r.returncode
    return r.stdout, r.stderr, sb


class Cosim:
    """Line-protocol client of oracle/_ref/ref_cosim (the reference's own code)."""

    def __init__(self, env=None, sb=None):
        build_all()
        self.sb = sb or Sandbox()
        e = preload_env(None, self.sb, env, stub=False)
        self.p = subprocess.Popen([REF_COSIM], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                  text=True, env=e, bufsize=1)

    def ask(self, line):
        self.p.stdin.write(line + "\n")
        self.p.stdin.flush()
        out = self.p.stdout.readline()
        if not out:
            raise RuntimeError("ref_cosim died on: " + line)
        return out.strip()

    def close(self):
        try:
            self.p.stdin.close()
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.sb.cleanup()
