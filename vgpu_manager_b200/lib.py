"""ctypes binding of the direct C ABI (include/vgpu_b200.h).

There is no Python or CPU fallback for any of these calls: if the shared object (with its
embedded sm_100a image) is missing, or the device runtime cannot be brought up in the current
CUDA context, `LibraryMissing` / `RuntimeError` is raised.
"""
import ctypes as C
import os
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SO = os.path.join(HERE, "libvgpu-control.so")


class LibraryMissing(RuntimeError):
    pass


class LimiterState(C.Structure):
    _fields_ = [("granted", C.c_longlong), ("consumed", C.c_longlong), ("bucket", C.c_longlong),
                ("share", C.c_longlong), ("up_limit", C.c_int), ("sys_free", C.c_int), ("avg_sys_free", C.c_int),
                ("ctr_i", C.c_int), ("pre_sys_process_num", C.c_int), ("valid", C.c_int),
                ("user_current", C.c_int), ("sys_current", C.c_int), ("sm_active_pct", C.c_int),
                ("queue_busy_pct", C.c_int), ("steps", C.c_ulonglong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class B200Library:
    """The interception library loaded into this process (not preloaded: hooks are not active,
    the direct entry points are).  `env` is applied before the first call because the library
    reads its contract from the environment exactly like the reference does."""

    def __init__(self, path=None, env=None, sandbox=True):
        path = path or DEFAULT_SO
        if not os.path.exists(path):
            raise LibraryMissing("%s not found - run `python -m vgpu_manager_b200.build` (needs nvcc for the "
                                 "sm_100a image); there is no CPU fallback" % path)
        if sandbox and "VGPU_B200_SANDBOX" not in os.environ:
            sb = tempfile.mkdtemp(prefix="vgpu_b200_sb_")
            for d in ("etc/vgpu-manager/config", "tmp/.vgpu_lock", "tmp/.vmem_node"):
                os.makedirs(os.path.join(sb, d), exist_ok=True)
            os.environ["VGPU_B200_SANDBOX"] = sb
        for k, v in (env or {}).items():
            os.environ[k] = v
        self.path = path
        self.h = C.CDLL(path)
        h = self.h
        h.vgpu_b200_version.restype = C.c_char_p
        h.vgpu_b200_clear.argtypes = [C.c_ulonglong, C.c_size_t, C.c_void_p]
        h.vgpu_b200_spill_copy.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_size_t, C.c_void_p]
        h.vgpu_b200_quota_eval.argtypes = [C.c_void_p, C.c_void_p]
        h.vgpu_b200_slab_insert.argtypes = [C.c_ulonglong, C.c_ulonglong]
        h.vgpu_b200_slab_remove.argtypes = [C.c_ulonglong, C.POINTER(C.c_ulonglong)]
        h.vgpu_b200_limiter_reset.argtypes = [C.c_int] * 6
        h.vgpu_b200_limiter_step.argtypes = [C.c_int] * 4 + [C.POINTER(LimiterState)]
        h.vgpu_b200_limiter_consume.argtypes = [C.c_longlong]
        h.vgpu_b200_refill.argtypes = [C.c_void_p, C.POINTER(LimiterState)]
        h.vgpu_b200_vslab_op.argtypes = [C.c_void_p, C.c_void_p]
        h.vgpu_b200_limiter_state.argtypes = [C.POINTER(LimiterState)]
        h.vgpu_b200_sampler_run.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_int, C.POINTER(LimiterState)]
        h.vgpu_b200_self_bytes.restype = C.c_ulonglong
        h.vgpu_b200_metric.restype = C.c_ulonglong
        h.vgpu_b200_metric.argtypes = [C.c_int, C.c_int]
        self._attached = False

    def version(self):
        return self.h.vgpu_b200_version().decode()

    def attach(self):
        """Bring up the device runtime in the calling thread's current CUDA context."""
        if self.h.vgpu_b200_attach() != 0:
            raise RuntimeError("vgpu_b200_attach failed: no current CUDA context or the sm_100a image could not "
                               "be loaded (see stderr); there is no CPU fallback")
        self._attached = True
        return self

    @staticmethod
    def _check(rc, what):
        if rc != 0:
            raise RuntimeError("%s failed with CUresult/rc %d" % (what, rc))

    def clear(self, ptr, nbytes, stream=0):
        self._check(self.h.vgpu_b200_clear(ptr, nbytes, C.c_void_p(stream)), "vgpu_b200_clear")

    def spill_copy(self, dst, src, nbytes, stream=0):
        self._check(self.h.vgpu_b200_spill_copy(dst, src, nbytes, C.c_void_p(stream)), "vgpu_b200_spill_copy")

    def quota_eval(self, req, res):
        self._check(self.h.vgpu_b200_quota_eval(C.byref(req), C.byref(res)), "vgpu_b200_quota_eval")
        return res

    def slab_insert(self, dptr, nbytes):
        return self.h.vgpu_b200_slab_insert(dptr, nbytes)

    def slab_remove(self, dptr):
        out = C.c_ulonglong(0)
        rc = self.h.vgpu_b200_slab_remove(dptr, C.byref(out))
        return rc, out.value

    def limiter_reset(self, sm_num, max_thread_per_sm, hard_core, soft_core, core_limit, hard_limit):
        self._check(self.h.vgpu_b200_limiter_reset(sm_num, max_thread_per_sm, hard_core, soft_core, core_limit,
                                                   hard_limit), "vgpu_b200_limiter_reset")

    def limiter_step(self, user, sys_, valid, nproc):
        st = LimiterState()
        self._check(self.h.vgpu_b200_limiter_step(user, sys_, valid, nproc, C.byref(st)), "vgpu_b200_limiter_step")
        return st

    def refill(self, util_req):
        """One default control step: device fold of the published samples + controller."""
        st = LimiterState()
        self._check(self.h.vgpu_b200_refill(C.byref(util_req), C.byref(st)), "vgpu_b200_refill")
        return st

    def vslab_op(self, req, res):
        """One slab-placement-table operation (structs mirror kernel_abi.h)."""
        self._check(self.h.vgpu_b200_vslab_op(C.byref(req), C.byref(res)), "vgpu_b200_vslab_op")
        return res

    def limiter_consume(self, tokens):
        self._check(self.h.vgpu_b200_limiter_consume(tokens), "vgpu_b200_limiter_consume")

    def limiter_state(self):
        st = LimiterState()
        self._check(self.h.vgpu_b200_limiter_state(C.byref(st)), "vgpu_b200_limiter_state")
        return st

    def sampler_run(self, window_us, interval_us, period_ticks, user_override=-1):
        st = LimiterState()
        self._check(self.h.vgpu_b200_sampler_run(window_us, interval_us, period_ticks, user_override, C.byref(st)),
                    "vgpu_b200_sampler_run")
        return st

    def set_spill_geometry(self, chunk, stages, ctas_per_sm):
        self._check(self.h.vgpu_b200_set_spill_geometry(chunk, stages, ctas_per_sm), "vgpu_b200_set_spill_geometry")

    def self_bytes(self):
        return self.h.vgpu_b200_self_bytes()
