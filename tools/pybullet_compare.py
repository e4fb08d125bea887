#!/usr/bin/env python
"""Pin the oracle against real PyBullet — runs ONLY where `pybullet`, `pybullet_data` and `gym` import.

Neither is installable in the build image or on the GPU box, so today this prints
"skipped".  Where they exist it runs the UNMODIFIED reference classes from a reference checkout
(`--reference /path/to/pybullet-robot-envs`) next to this repo's CPU oracle on the same seed and
action stream and reports the divergence per step: joint angles, EE pose, cube pose, reward, done,
and the contact count (p.getContactPoints).  The first N steps before any robot contact are the
meaningful ones (contact dynamics are chaotic and the link geometry differs: sphere proxies vs meshes).

    python tools/pybullet_compare.py --reference /root/reference --steps 240 --seed 0
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    try:
        import pybullet  # noqa: F401
        import pybullet_data  # noqa: F401
        import gym  # noqa: F401
    except Exception as e:  # the expected outcome in this image
        print("skipped: pybullet / pybullet_data / gym not importable (%s) — parity with PyBullet stays unpinned" % e)
        return 0
    import numpy as np
    sys.path.insert(0, args.reference)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
    time.sleep = lambda *_: None                      # the reference sleeps 1/240 s per step (quirk E.1)
    from pybullet_robot_envs.envs.panda_envs.panda_push_gym_env import pandaPushGymEnv as RefEnv  # reference class
    from oracle import b2oracle
    from pybullet_robot_envs.b2env.model import TASK_PUSH, panda_task_setup

    ref = RefEnv(renders=False, obj_pose_rnd_std=0.05, tg_pose_rnd_std=0, max_steps=1000)
    ref.seed(args.seed)
    obs_ref = ref.reset()
    m, p = panda_task_setup(TASK_PUSH)
    orc = b2oracle.Oracle(m, p, 1)
    # start the oracle from the reference's post-reset state so that only the step is compared
    import pybullet as pb
    cid = ref._physics_client_id
    q = [pb.getJointState(ref._robot.robot_id, j, physicsClientId=cid)[0] for j in ref._robot._joint_name_to_ids.values()]
    pos, orn = pb.getBasePositionAndOrientation(ref._world.obj_id, physicsClientId=cid)
    pose = np.array([list(pos) + list(orn)], np.float32)
    orc.reset(pose, np.array([ref._target_pose], np.float32))
    orc.state["q"][0] = q
    orc.state["mtarget"][0] = q
    rng = np.random.RandomState(1234)
    print("step  max|dq|    |dEE|     |dcube|   dreward  done(ref,oracle)  contacts(ref)")
    for t in range(args.steps):
        a = rng.uniform(-1, 1, 7).astype(np.float32)
        o_r, r_r, d_r, _ = ref.step(a)
        o_o, r_o, d_o = orc.step(a[None, :], 1, 0)
        q_r = np.array([pb.getJointState(ref._robot.robot_id, j, physicsClientId=cid)[0] for j in ref._robot._joint_name_to_ids.values()])
        pos, _ = pb.getBasePositionAndOrientation(ref._world.obj_id, physicsClientId=cid)
        raw_r = ref.get_extended_observation()[0]
        nct = len(pb.getContactPoints(ref._world.obj_id, ref._robot.robot_id, physicsClientId=cid))
        print("%4d  %.2e  %.2e  %.2e  %+.2e  %d %d  %d" % (
            t, np.abs(q_r - orc.state["q"][0]).max(), np.linalg.norm(raw_r[:3] - orc.state["raw_obs"][0, :3]),
            np.linalg.norm(np.array(pos) - orc.state["obj_pose"][0, :3]), float(r_r) - float(r_o[0]), int(d_r), int(d_o[0]), nct))
    return 0


if __name__ == "__main__":
    sys.exit(main())
