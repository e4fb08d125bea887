"""``pandaPushGymEnv`` — batched, CUDA-backed counterpart of reference
envs/panda_envs/panda_push_gym_env.py:24-376 (same constructor kwargs + ``num_envs``/``device``)."""
import numpy as np

from pybullet_robot_envs.b2env.model import TASK_PUSH
from pybullet_robot_envs.envs.world_envs.world_env import get_objects_list
from pybullet_robot_envs.envs.panda_envs._panda_task import PandaTaskBase


class pandaPushGymEnv(PandaTaskBase):
    _task = TASK_PUSH
    _is_task_impl = True

    def __init__(self, numControlledJoints=7, use_IK=0, action_repeat=1, obj_name=get_objects_list()[1],
                 renders=False, max_steps=1000, obj_pose_rnd_std=0.0, tg_pose_rnd_std=0.0, includeVelObs=True,
                 num_envs=1, device=0):
        self._target_dist_max = 0.3
        self._tg_pose_rnd_std = tg_pose_rnd_std
        # success radius 0.1 (reference :52); robot workspace floor = table - 0.2 (:74)
        self._setup(numControlledJoints, use_IK, action_repeat, obj_name, renders, max_steps, obj_pose_rnd_std,
                    includeVelObs, num_envs, device, target_dist_min=0.1, z_low_offset=-0.2)

    def sample_tg_pose(self, obj_pos):
        """Target = object + (0.05, 0.05, 0), or object + N(0, tg_pose_rnd_std) drawn from the
        GLOBAL numpy generator like the reference (:333-360, quirk E.10); clipped to the workspace."""
        obj_pos = np.asarray(obj_pos, np.float64).reshape(-1, 3)
        ws = self._world.get_workspace()
        x_min, x_max = ws[0][0] + 0.07, ws[0][1] - 0.07
        y_min, y_max = ws[1][0], ws[1][1]
        px = obj_pos[:, 0] + 0.05
        py = obj_pos[:, 1] + 0.05
        pz = obj_pos[:, 2]
        if self._tg_pose_rnd_std > 0:
            noise = np.random.normal(0, self._tg_pose_rnd_std, (obj_pos.shape[0], 2))
            px = obj_pos[:, 0] + noise[:, 0]
            py = obj_pos[:, 1] + noise[:, 1]
        px = np.clip(px, x_min, x_max)
        py = np.clip(py, y_min, y_max)
        pose = np.stack([px, py, pz], axis=1)
        return tuple(pose[0]) if self.num_envs == 1 else pose
