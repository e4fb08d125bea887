"""ctypes binding of libb2env.so (include/b2env.h) — the only door to the CUDA path.

There is no CPU fallback: if the library is missing or no CUDA device is present the
constructor raises, so a silent non-GPU path cannot exist.
"""
import ctypes as C
import os

import numpy as np

from .model import B2EModel, B2EParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2ENV_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "csrc", "libb2env.so"))

F_Q, F_QD, F_OBJ_POSE, F_OBJ_VEL, F_TARGET, F_MTARGET, F_COUNTERS, F_CACHE_KEY, F_CACHE_LAM, \
    F_HAND_POSE, F_STATUS, F_RAW_OBS, F_CONTACTS, F_SHAPING = range(14)
FIELD_NAMES = {
    "q": F_Q, "qd": F_QD, "obj_pose": F_OBJ_POSE, "obj_vel": F_OBJ_VEL, "target": F_TARGET,
    "mtarget": F_MTARGET, "counters": F_COUNTERS, "cache_key": F_CACHE_KEY, "cache_lam": F_CACHE_LAM,
    "hand_pose": F_HAND_POSE, "status": F_STATUS, "raw_obs": F_RAW_OBS, "contacts": F_CONTACTS,
    "shaping": F_SHAPING,
}
INT_FIELDS = {F_COUNTERS, F_CACHE_KEY, F_STATUS}
MODE_ACTION, MODE_HOLD, MODE_TARGETS, MODE_IK_POSE, MODE_OBSERVE = 0, 1, 2, 3, 4
OPT_RECORD_CONTACTS = 0

EXPORTS = [
    "b2e_create", "b2e_destroy", "b2e_set_params", "b2e_set_option", "b2e_reset", "b2e_step",
    "b2e_step_subset", "b2e_set_rows", "b2e_get_rows",
    "b2e_step_host", "b2e_step_pinned", "b2e_host_alloc", "b2e_host_free", "b2e_get", "b2e_set", "b2e_get_host", "b2e_set_host", "b2e_field_width",
    "b2e_field_elem_size", "b2e_num_envs", "b2e_launch_count", "b2e_timer_start", "b2e_timer_stop",
    "b2e_last_error", "b2e_version", "b2e_debug_sched",
]

_lib = None


class B2EError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libb2env.so and declare prototypes.  Raises if the extension is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise B2EError("CUDA extension %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    vp, ci = C.c_void_p, C.c_int
    lib.b2e_create.argtypes = [C.POINTER(B2EModel), C.POINTER(B2EParams), ci, ci, C.POINTER(vp)]
    lib.b2e_destroy.argtypes = [vp]
    lib.b2e_destroy.restype = None
    lib.b2e_set_params.argtypes = [vp, C.POINTER(B2EParams)]
    lib.b2e_set_option.argtypes = [vp, ci, ci]
    lib.b2e_reset.argtypes = [vp, vp, vp, vp, vp]
    lib.b2e_step.argtypes = [vp, vp, vp, vp, vp, ci, ci, vp]
    lib.b2e_step_subset.argtypes = [vp, vp, ci, vp, vp, vp, vp, ci, ci, vp]
    lib.b2e_set_rows.argtypes = [vp, ci, vp, ci, vp, vp]
    lib.b2e_get_rows.argtypes = [vp, ci, vp, ci, vp, vp]
    lib.b2e_step_host.argtypes = [vp, vp, vp, vp, vp, ci, ci]
    lib.b2e_step_pinned.argtypes = [vp, vp, vp, vp, vp, ci, ci]
    lib.b2e_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.b2e_host_free.argtypes = [vp]
    lib.b2e_get.argtypes = [vp, ci, vp, vp]
    lib.b2e_set.argtypes = [vp, ci, vp, vp]
    lib.b2e_get_host.argtypes = [vp, ci, vp]
    lib.b2e_set_host.argtypes = [vp, ci, vp]
    lib.b2e_field_width.argtypes = [vp, ci]
    lib.b2e_field_elem_size.argtypes = [ci]
    lib.b2e_num_envs.argtypes = [vp]
    lib.b2e_launch_count.argtypes = [vp]
    lib.b2e_launch_count.restype = C.c_int64
    lib.b2e_timer_start.argtypes = [vp, vp]
    lib.b2e_timer_stop.argtypes = [vp, vp, C.POINTER(C.c_float)]
    lib.b2e_debug_sched.argtypes = [vp, vp, ci, C.POINTER(ci), C.POINTER(ci)]
    lib.b2e_last_error.restype = C.c_char_p
    lib.b2e_version.restype = C.c_char_p
    if path is None:
        _lib = lib
    return lib


def _ptr(x):
    """Device/host pointer of a torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(x.data_ptr())  # torch tensor


class B2Sim:
    """Thin object wrapper over the C-ABI: one simulation of ``num_envs`` envs on one GPU."""

    def __init__(self, model: B2EModel, params: B2EParams, num_envs: int, device: int = 0, lib=None):
        # `lib`: tests may hand in another build of the SAME source (tools/emu: the kernels compiled for the
        # host, to debug kernel logic without a GPU).  The product never passes it.
        self.lib = lib if lib is not None else load_library()
        # the emulation build (tests only) keeps "device" memory on the host: its device pointers are numpy arrays
        self._host_mem = b"EMULATION" in self.lib.b2e_version()
        self.model, self.params, self.B, self.device = model, params, int(num_envs), int(device)
        h = C.c_void_p()
        self._check(self.lib.b2e_create(C.byref(model), C.byref(params), self.B, self.device, C.byref(h)))
        self.h = h
        self._pin = None
        self._pinned_ptrs = []
        self._pinned_ranges = []

    def _check(self, rc):
        if rc != 0:
            raise B2EError("libb2env error %d: %s" % (rc, self.lib.b2e_last_error().decode()))

    def close(self):
        if getattr(self, "h", None):
            self._pin = None
            for p in getattr(self, "_pinned_ptrs", []):
                self.lib.b2e_host_free(p)
            self._pinned_ptrs = []
            self._pinned_ranges = []
            self.lib.b2e_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        self._check(self.lib.b2e_set_params(self.h, C.byref(params)))
        self.params = params

    def set_option(self, option, value):
        self._check(self.lib.b2e_set_option(self.h, option, int(value)))

    # device-pointer entry points (torch tensors on the sim's device, or raw ints)
    def reset(self, obj_init_pose, target, env_mask=None, stream=0):
        self._check(self.lib.b2e_reset(self.h, _ptr(env_mask), _ptr(obj_init_pose), _ptr(target), C.c_void_p(stream)))

    def step(self, action, obs=None, reward=None, done=None, n_substeps=1, mode=MODE_ACTION, stream=0):
        self._check(self.lib.b2e_step(self.h, _ptr(action), _ptr(obs), _ptr(reward), _ptr(done), n_substeps, mode,
                                      C.c_void_p(stream)))

    # page-locked numpy buffers owned by the library
    def _pinned(self, shape, dtype=np.float32):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._check(self.lib.b2e_host_alloc(C.byref(p), n))
        self._pinned_ptrs.append(p)
        self._pinned_ranges.append((p.value, p.value + n))
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def pinned_array(self, shape, dtype=np.float32):
        """Page-locked numpy array owned by this simulation (freed by close()): action batches kept in such
        arrays are copied to the GPU straight from user memory by step_pinned."""
        return self._pinned(shape, dtype)

    def _is_pinned(self, arr):
        a = arr.ctypes.data
        return arr.flags["C_CONTIGUOUS"] and any(lo <= a and a + arr.nbytes <= hi for lo, hi in self._pinned_ranges)

    def _ensure_pin(self):
        if self._pin is None:
            self._pin = (self._pinned((self.B, self.params.n_act)), self._pinned((self.B, self.params.n_obs)),
                         self._pinned((self.B,)), self._pinned((self.B,)))
        return self._pin

    def step_pinned(self, action, n_substeps=1, mode=MODE_ACTION):
        """Fast host path: `action` is copied into a page-locked staging array (or IS the array returned by
        `pinned_action()`), results land in page-locked arrays that are returned as views — valid until the
        next call."""
        pa, po, pr, pd = self._ensure_pin()
        src = pa
        if action is not pa:
            act = action if isinstance(action, np.ndarray) and action.dtype == np.float32 else np.asarray(action, dtype=np.float32)
            if act.shape != pa.shape:
                raise AssertionError(("number of motor commands differs from number of motor to control", act.shape))
            if self._is_pinned(act):
                src = act                      # already page-locked: no staging copy
            else:
                np.copyto(pa, act)
        self._check(self.lib.b2e_step_pinned(self.h, _ptr(src), _ptr(po), _ptr(pr), _ptr(pd), n_substeps, mode))
        return po, pr, pd

    def pinned_action(self):
        """Page-locked [B, n_act] array a policy can write into directly (skips one host copy)."""
        return self._ensure_pin()[0]

    # host-buffer entry points (numpy)
    def step_host(self, action=None, n_substeps=1, mode=MODE_ACTION, want_obs=True):
        no = self.params.n_obs
        act = None
        if action is not None:
            act = np.ascontiguousarray(action, dtype=np.float32)
            if act.shape != (self.B, self.params.n_act):
                raise AssertionError(("number of motor commands differs from number of motor to control", act.shape))
        if not want_obs:
            self._check(self.lib.b2e_step_host(self.h, _ptr(act), None, None, None, n_substeps, mode))
            return None
        obs = np.empty((self.B, no), np.float32)
        rew = np.empty(self.B, np.float32)
        done = np.empty(self.B, np.float32)
        self._check(self.lib.b2e_step_host(self.h, _ptr(act), _ptr(obs), _ptr(rew), _ptr(done), n_substeps, mode))
        return obs, rew, done

    def get(self, name):
        f = FIELD_NAMES[name]
        w = self.lib.b2e_field_width(self.h, f)
        out = np.empty((self.B, w), np.int32 if f in INT_FIELDS else np.float32)
        self._check(self.lib.b2e_get_host(self.h, f, _ptr(out)))
        return out

    def get_device(self, name, like):
        """Device-side copy of a state field into a new torch tensor on ``like``'s device / current stream."""
        import torch
        f = FIELD_NAMES[name]
        w = self.lib.b2e_field_width(self.h, f)
        out = torch.empty((self.B, w), device=like.device, dtype=torch.int32 if f in INT_FIELDS else torch.float32)
        self._check(self.lib.b2e_get(self.h, f, _ptr(out), C.c_void_p(torch.cuda.current_stream(like.device).cuda_stream)))
        return out

    def set(self, name, value):
        f = FIELD_NAMES[name]
        w = self.lib.b2e_field_width(self.h, f)
        arr = np.ascontiguousarray(value, np.int32 if f in INT_FIELDS else np.float32).reshape(self.B, w)
        self._check(self.lib.b2e_set_host(self.h, f, _ptr(arr)))

    def reset_host(self, obj_init_pose, target, env_mask=None):
        """Reset from host arrays (copies through temporary device buffers owned by torch)."""
        o = self._to_dev(obj_init_pose, np.float32)
        t = self._to_dev(target, np.float32)
        m = None if env_mask is None else self._to_dev(env_mask, np.uint8)
        self._sync()
        self.reset(o, t, m)
        self._sync()

    # subset entry points (per-env resets)
    def _to_dev(self, arr, dtype):
        a = np.ascontiguousarray(arr, dtype)
        if self._host_mem:
            return a
        import torch
        return torch.as_tensor(a).to(torch.device("cuda", self.device))

    def _sync(self):
        if not self._host_mem:
            import torch
            torch.cuda.synchronize(torch.device("cuda", self.device))

    def _ids_dev(self, ids):
        return self._to_dev(ids, np.int32)

    def step_subset(self, ids, n_substeps=1, mode=MODE_HOLD):
        """Advance only the listed environments (no action / outputs: settle steps of a reset)."""
        d = self._ids_dev(ids)
        self._check(self.lib.b2e_step_subset(self.h, _ptr(d), int(len(d)), None, None, None, None, n_substeps, mode,
                                             C.c_void_p(0)))
        self._sync()

    def set_rows(self, name, ids, values):
        f = FIELD_NAMES[name]
        w = self.lib.b2e_field_width(self.h, f)
        d = self._ids_dev(ids)
        v = self._to_dev(np.asarray(values).reshape(-1, w), np.int32 if f in INT_FIELDS else np.float32)
        assert v.shape[0] == len(d)
        self._check(self.lib.b2e_set_rows(self.h, f, _ptr(d), int(len(d)), _ptr(v), C.c_void_p(0)))
        self._sync()

    def get_rows(self, name, ids):
        f = FIELD_NAMES[name]
        w = self.lib.b2e_field_width(self.h, f)
        d = self._ids_dev(ids)
        out = self._to_dev(np.zeros((len(d), w)), np.int32 if f in INT_FIELDS else np.float32)
        self._check(self.lib.b2e_get_rows(self.h, f, _ptr(d), int(len(d)), _ptr(out), C.c_void_p(0)))
        self._sync()
        return out if self._host_mem else out.cpu().numpy()

    def debug_sched_lists(self):
        """(environments in scheduling order, how many of them are in the tail list) — diagnostics for the tests."""
        out = np.zeros(self.B, np.int32)
        nm, nt = C.c_int(), C.c_int()
        self._check(self.lib.b2e_debug_sched(self.h, _ptr(out), self.B, C.byref(nm), C.byref(nt)))
        return [int(x) for x in out[:nm.value + nt.value]], nt.value

    def launch_count(self):
        return int(self.lib.b2e_launch_count(self.h))

    def timer_start(self, stream=0):
        self._check(self.lib.b2e_timer_start(self.h, C.c_void_p(stream)))

    def timer_stop(self, stream=0):
        ms = C.c_float()
        self._check(self.lib.b2e_timer_stop(self.h, C.c_void_p(stream), C.byref(ms)))
        return ms.value
