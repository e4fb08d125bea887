"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol declared in
include/b2env.h, struct mirrors have the C layout, gym shim, utils, the missing-GPU failure mode."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from pybullet_robot_envs.b2env import binding
from pybullet_robot_envs.b2env.model import B2EModel, B2EParams, TASK_PUSH, panda_task_setup
from pybullet_robot_envs.gym_compat import spaces
from pybullet_robot_envs.envs.utils import goal_distance, scale_gym_data, unscale_gym_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b2env.h")).read()
    declared = sorted(set(re.findall(r"\b(b2e_[a-z_]+)\s*\(", hdr)))
    assert len(declared) >= 19
    lib = binding.load_library()
    for name in declared:
        assert hasattr(lib, name), "libb2env.so does not export %s" % name
    assert sorted(binding.EXPORTS) == declared
    assert b"sm_100a" in lib.b2e_version()


def test_struct_layouts_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b2env.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(b2e_model),sizeof(b2e_params),offsetof(b2e_model,sph_link),offsetof(b2e_params,obs_low));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c, d = (int(x) for x in subprocess.check_output([str(exe)]).split())
    assert a == C.sizeof(B2EModel) and b == C.sizeof(B2EParams)
    assert c == B2EModel.sph_link.offset and d == B2EParams.obs_low.offset


def test_no_gpu_fails_loudly():
    """In the CPU container the product path must refuse to run (no fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m, p = panda_task_setup(TASK_PUSH)
    with pytest.raises(binding.B2EError, match="no CPU fallback"):
        binding.B2Sim(m, p, 4, 0)
    import pybullet_robot_envs
    from pybullet_robot_envs import gym
    with pytest.raises(binding.B2EError):
        gym.make("pandaPush-v0")


def test_registered_ids():
    import pybullet_robot_envs
    from pybullet_robot_envs.gym_compat import registry
    assert registry.spec("PandaGrasp-v0").max_episode_steps == 1000 and registry.spec("pandaPushGoal-v0")
    for i in ("pandaReach-v0", "pandaPush-v0", "PandaReach-v0", "PandaPush-v0"):
        s = registry.spec(i)
        assert s.max_episode_steps == 1000
        assert s._kwargs["use_IK"] == 0 and s._kwargs["obj_pose_rnd_std"] == 0.05
    assert registry.spec("pandaPush-v0")._kwargs["tg_pose_rnd_std"] == 0


def test_utils_scale_and_distance():
    sp = spaces.Box(np.array([0.0, -2.0]), np.array([1.0, 2.0]), dtype="float32")
    np.testing.assert_allclose(scale_gym_data(sp, np.array([0.5, 0.0])), [0, 0])
    np.testing.assert_allclose(scale_gym_data(sp, np.array([2.0, 4.0])), [3, 2])        # no clipping
    x = np.array([[0.1, 1.0], [0.9, -1.0]])
    np.testing.assert_allclose(unscale_gym_data(sp, scale_gym_data(sp, x)), x, atol=1e-7)
    with pytest.raises(AssertionError):
        scale_gym_data(sp, np.zeros(3))
    assert goal_distance(np.array([0, 0, 0.0]), np.array([3, 4, 0.0])) == 5.0
    np.testing.assert_allclose(goal_distance(np.zeros((4, 3)), np.ones((4, 3))), np.sqrt(3) * np.ones(4))
    with pytest.raises(AssertionError):
        goal_distance(np.zeros(3), np.zeros(2))


def test_box_space():
    b = spaces.Box(-np.ones(7), np.ones(7), dtype="float32")
    assert b.shape == (7,) and b.low.dtype == np.float32
    b.seed(0)
    s = b.sample()
    assert s.shape == (7,) and b.contains(s)
