"""Collision proxies for the Panda links.

The reference collides convex hulls of ``robot_data/franka_panda/meshes/collision/*.obj``
(reference panda_model.urdf, ``<collision>`` tags), but those meshes are git-LFS pointer
stubs in the reference tree, so the geometry is unavailable.  The links
that can reach the table or the object are covered by spheres sized from the publicly
known Panda dimensions.  Centres are in the owning link's URDF frame, metres.
Every parity claim about robot-link contacts is therefore proxy-based.
"""

PANDA_SPHERES = [
    dict(link="panda_link4", c=(-0.04, 0.04, 0.0), r=0.06),
    dict(link="panda_link5", c=(0.0, 0.06, -0.20), r=0.055),
    dict(link="panda_link5", c=(0.0, 0.08, -0.08), r=0.05),
    dict(link="panda_link6", c=(0.04, 0.0, 0.0), r=0.06),
    dict(link="panda_link7", c=(0.0, 0.0, 0.07), r=0.05),
    dict(link="panda_hand", c=(0.0, 0.0, 0.03), r=0.035),
    dict(link="panda_hand", c=(0.0, 0.06, 0.03), r=0.03),
    dict(link="panda_hand", c=(0.0, -0.06, 0.03), r=0.03),
    dict(link="panda_leftfinger", c=(0.0, 0.008, 0.045), r=0.009),
    dict(link="panda_leftfinger", c=(0.0, 0.008, 0.027), r=0.009),
    dict(link="panda_leftfinger", c=(0.0, 0.008, 0.009), r=0.009),
    dict(link="panda_rightfinger", c=(0.0, -0.008, 0.045), r=0.009),
    dict(link="panda_rightfinger", c=(0.0, -0.008, 0.027), r=0.009),
    dict(link="panda_rightfinger", c=(0.0, -0.008, 0.009), r=0.009),
]

# contact feature keys (stable across steps; used by the warm-start cache)
KEY_CUBE_TABLE = 0    # + cube vertex index 0..7
KEY_CUBE_PLANE = 8    # + cube vertex index 0..7
KEY_SPHERE_CUBE = 16  # + sphere index
KEY_SPHERE_TABLE = 32 # + sphere index
