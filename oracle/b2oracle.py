"""ctypes wrapper of the CPU oracle (oracle/b2oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, os.path.join(_ROOT, "pybullet-robot-envs_b200"))
from pybullet_robot_envs.b2env.model import (B2EModel, B2EParams, CACHE_SLOTS, MAX_CONTACTS)  # noqa: E402


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc; a few seconds)."""
    so = os.path.join(_HERE, "libb2oracle.so")
    src = os.path.join(_HERE, "b2oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


class _State(C.Structure):
    _fields_ = [("B", C.c_int32)] + [(n, C.c_void_p) for n in (
        "q", "qd", "obj_pose", "obj_vel", "target", "mtarget", "counters", "cache_key", "cache_lam",
        "hand_pose", "status", "raw_obs", "contacts", "shaping")]


FIELDS = {  # name -> (width fn(n_dof, n_obs), dtype)
    "q": (lambda nd, no: nd, np.float32), "qd": (lambda nd, no: nd, np.float32),
    "obj_pose": (lambda nd, no: 7, np.float32), "obj_vel": (lambda nd, no: 6, np.float32),
    "target": (lambda nd, no: 3, np.float32), "mtarget": (lambda nd, no: nd, np.float32),
    "counters": (lambda nd, no: 2, np.int32), "cache_key": (lambda nd, no: CACHE_SLOTS, np.int32),
    "cache_lam": (lambda nd, no: CACHE_SLOTS * 3, np.float32), "hand_pose": (lambda nd, no: 6, np.float32),
    "status": (lambda nd, no: 4, np.int32), "raw_obs": (lambda nd, no: no, np.float32),
    "contacts": (lambda nd, no: MAX_CONTACTS * 8, np.float32),
    "shaping": (lambda nd, no: 2, np.float32),
}


class Oracle:
    """Batched CPU simulation with the same state layout as the CUDA library."""

    def __init__(self, model: B2EModel, params: B2EParams, num_envs: int, double=False, nthreads=1):
        build()
        name = "libb2oracle_f64.so" if double else "libb2oracle.so"
        self.lib = C.CDLL(os.path.join(_HERE, name))
        self.model, self.params, self.B, self.nthreads = model, params, num_envs, nthreads
        nd, no = model.n_dof, params.n_obs
        self.state = {k: np.zeros((num_envs, w(nd, no)), dtype=dt) for k, (w, dt) in FIELDS.items()}
        self.state["cache_key"][:] = -1
        self._st = _State()
        self._st.B = num_envs
        for k, a in self.state.items():
            setattr(self._st, k, a.ctypes.data)
        self.lib.b2o_step.restype = C.c_int
        self.lib.b2o_reset.restype = C.c_int

    def reset(self, obj_init_pose, target, mask=None):
        obj = np.ascontiguousarray(obj_init_pose, dtype=np.float32)
        tg = np.ascontiguousarray(target, dtype=np.float32)
        mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = self.lib.b2o_reset(C.byref(self.model), C.byref(self.params), C.byref(self._st),
                                None if mk is None else mk.ctypes.data_as(C.c_void_p),
                                obj.ctypes.data_as(C.c_void_p), tg.ctypes.data_as(C.c_void_p))
        assert rc == 0

    def step(self, action=None, n_substeps=1, mode=0, want_obs=True):
        no = self.params.n_obs
        obs = np.zeros((self.B, no), np.float32)
        rew = np.zeros(self.B, np.float32)
        done = np.zeros(self.B, np.float32)
        act = None
        if action is not None:
            act = np.ascontiguousarray(action, dtype=np.float32)
            assert act.shape == (self.B, self.params.n_act), act.shape
        rc = self.lib.b2o_step(C.byref(self.model), C.byref(self.params), C.byref(self._st),
                               None if act is None else act.ctypes.data_as(C.c_void_p),
                               obs.ctypes.data_as(C.c_void_p) if want_obs else None,
                               rew.ctypes.data_as(C.c_void_p) if want_obs else None,
                               done.ctypes.data_as(C.c_void_p) if want_obs else None,
                               C.c_int(n_substeps), C.c_int(mode), C.c_int(self.nthreads))
        assert rc == 0
        return obs, rew, done

    # ---- known-answer helpers -------------------------------------------------
    def fk(self, q):
        n = self.model.n_links
        pos = np.zeros((n, 3), np.float32)
        rot = np.zeros((n, 9), np.float32)
        q = np.ascontiguousarray(q, np.float32)
        self.lib.b2o_fk(C.byref(self.model), q.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p),
                        rot.ctypes.data_as(C.c_void_p))
        return pos, rot.reshape(n, 3, 3)

    def forward_dynamics(self, q, qd, tau):
        nd = self.model.n_dof
        q, qd, tau = (np.ascontiguousarray(x, np.float32) for x in (q, qd, tau))
        out = np.zeros(nd, np.float32)
        self.lib.b2o_forward_dynamics(C.byref(self.model), C.byref(self.params), q.ctypes.data_as(C.c_void_p),
                                      qd.ctypes.data_as(C.c_void_p), tau.ctypes.data_as(C.c_void_p),
                                      out.ctypes.data_as(C.c_void_p))
        return out

    def minv(self, q):
        nd = self.model.n_dof
        q = np.ascontiguousarray(q, np.float32)
        out = np.zeros((nd, nd), np.float32)
        self.lib.b2o_minv(C.byref(self.model), q.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def ee_jacobian(self, q):
        nd = self.model.n_dof
        q = np.ascontiguousarray(q, np.float32)
        J = np.zeros((6, nd), np.float32)
        pos = np.zeros(3, np.float32)
        quat = np.zeros(4, np.float32)
        self.lib.b2o_ee_jacobian(C.byref(self.model), q.ctypes.data_as(C.c_void_p), J.ctypes.data_as(C.c_void_p),
                                 pos.ctypes.data_as(C.c_void_p), quat.ctypes.data_as(C.c_void_p))
        return J, pos, quat
