#!/usr/bin/env python
"""Generate the committed Panda model descriptor from the reference URDF.

Run in the build container (where /root/reference exists):
    python tools/gen_model.py
It parses /root/reference/pybullet_robot_envs/robot_data/franka_panda/panda_model.urdf with
this repo's own URDF parser and writes the kinematic/inertial data (no meshes, no XML) to
pybullet-robot-envs_b200/pybullet_robot_envs/robot_data/franka_panda/panda_model.json,
because /root/reference does not exist on the GPU box.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
from pybullet_robot_envs.b2env.model import parse_urdf, parse_sdf, PANDA_JSON, ICUB_JSON  # noqa: E402

SRC = "/root/reference/pybullet_robot_envs/robot_data/franka_panda/panda_model.urdf"

if __name__ == "__main__":
    d = parse_urdf(SRC)
    d["source"] = "hsp-iit/pybullet-robot-envs robot_data/franka_panda/panda_model.urdf (Apache-2.0 data)"
    os.makedirs(os.path.dirname(PANDA_JSON), exist_ok=True)
    with open(PANDA_JSON, "w") as f:
        json.dump(d, f, indent=1)
    print("wrote", os.path.normpath(PANDA_JSON), len(d["joints"]), "joints")
    d = parse_sdf("/root/reference/pybullet_robot_envs/robot_data/iCub/icub_model.sdf")
    d["source"] = "hsp-iit/pybullet-robot-envs robot_data/iCub/icub_model.sdf (kinematic/inertial data only)"
    os.makedirs(os.path.dirname(ICUB_JSON), exist_ok=True)
    with open(ICUB_JSON, "w") as f:
        json.dump(d, f, indent=1)
    print("wrote", os.path.normpath(ICUB_JSON), len(d["joints"]), "joints")
