"""``iCubPushGymEnv`` — batched, CUDA-backed counterpart of reference
envs/icub_envs/icub_push_gym_env.py:23-410 (same constructor kwargs + ``num_envs``/``device``)."""
import numpy as np

from pybullet_robot_envs.b2env.model import TASK_PUSH
from pybullet_robot_envs.envs.icub_envs._icub_task import ICubTaskBase
from pybullet_robot_envs.envs.world_envs.world_env import get_objects_list


class iCubPushGymEnv(ICubTaskBase):
    _task = TASK_PUSH
    _is_task_impl = True

    def __init__(self, action_repeat=1, use_IK=1, control_arm='l', control_orientation=0, obj_name=get_objects_list()[1],
                 obj_pose_rnd_std=0, tg_pose_rnd_std=0.2, renders=False, max_steps=2000, reward_type=1, num_envs=1,
                 device=0):
        self._tg_pose_rnd_std = tg_pose_rnd_std
        self._reward_type = reward_type
        self._init_dist_hand_obj = None
        self._max_dist_obj_tg = None
        self._setup_icub(action_repeat, use_IK, control_arm, control_orientation, obj_name, obj_pose_rnd_std, renders,
                         max_steps, num_envs, device)

    def _after_target(self, ids=None):
        """The two distances the type-1 reward is normalised with, captured at reset (reference :120-127)."""
        raw = self._physics_client_id.observe()[3]
        tp = np.asarray(self._target_pose, np.float32).reshape(-1, 3)
        d0 = np.linalg.norm(raw[:, 0:3] - raw[:, 19:22], axis=1)
        dm = np.linalg.norm(raw[:, 19:22] - np.broadcast_to(tp, (self.num_envs, 3)), axis=1)
        sh = np.stack([d0, dm], axis=1).astype(np.float32)
        if ids is None:
            self._init_dist_hand_obj, self._max_dist_obj_tg = (d0[0], dm[0]) if self.num_envs == 1 else (d0, dm)
            self._physics_client_id.set("shaping", sh)
        else:
            self._physics_client_id.set_rows("shaping", ids, sh[ids])
            if self.num_envs == 1:
                self._init_dist_hand_obj, self._max_dist_obj_tg = d0[0], dm[0]
            else:   # keep the host-side attributes (read by subclasses that override _compute_reward) in step
                self._init_dist_hand_obj = np.array(np.broadcast_to(self._init_dist_hand_obj, (self.num_envs,)), np.float32)
                self._max_dist_obj_tg = np.array(np.broadcast_to(self._max_dist_obj_tg, (self.num_envs,)), np.float32)
                self._init_dist_hand_obj[ids] = d0[ids]
                self._max_dist_obj_tg[ids] = dm[ids]

    def sample_tg_pose(self, obj_pos):
        """Target = object + (0.05, 0.05, 0), or object + N(0, tg_pose_rnd_std) from the GLOBAL numpy generator
        like the reference (:375-398); clipped to the world workspace."""
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
        pose = np.stack([np.clip(px, x_min, x_max), np.clip(py, y_min, y_max), pz], axis=1)
        return tuple(pose[0]) if self.num_envs == 1 else pose
