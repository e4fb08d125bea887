#!/usr/bin/env python
"""Self-collision candidate pairs of the Panda sphere proxies (proxies.PANDA_SELF_PAIRS): sphere pairs on links at least
three joints apart whose gap can fall below the contact margin somewhere in the joint ranges (20000 random configurations,
forward kinematics of the oracle).  Prints the list to paste into proxies.py (at most 32 pairs, closest first)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pybullet-robot-envs_b200"))
from oracle import b2oracle  # noqa: E402
from pybullet_robot_envs.b2env.model import TASK_PUSH, panda_task_setup  # noqa: E402

m, p = panda_task_setup(TASK_PUSH)
o = b2oracle.Oracle(m, p, 1)
S = m.n_spheres
par = [m.parent[i] for i in range(m.n_links)]


def chain(x):
    out = [x]
    while par[x] >= 0:
        x = par[x]
        out.append(x)
    return out + [-1]


def kin_dist(a, b):
    A, B = chain(a), chain(b)
    for i, x in enumerate(A):
        if x in B:
            return i + B.index(x)


def centres(q):
    pos, rot = o.fk(q)
    return np.array([pos[m.sph_link[s]] + rot[m.sph_link[s]] @ np.array(m.sph_c[s][:]) for s in range(S)])


lo = np.array([m.lower[i] for i in range(9)])
hi = np.array([m.upper[i] for i in range(9)])
rng = np.random.RandomState(0)
mind = np.full((S, S), 9.0)
for t in range(20000):
    c = centres((lo + (hi - lo) * rng.uniform(0, 1, 9)).astype(np.float32))
    mind = np.minimum(mind, np.linalg.norm(c[:, None] - c[None], axis=2))
cand = []
for a in range(S):
    for b in range(a + 1, S):
        if kin_dist(m.sph_link[a], m.sph_link[b]) < 3:
            continue
        gap = mind[a, b] - (m.sph_r[a] + m.sph_r[b])
        if gap < p.contact_margin:
            cand.append((gap, a, b))
cand.sort()
print("PANDA_SELF_PAIRS = [" + ", ".join("(%d, %d)" % (a, b) for _, a, b in cand[:32]) + "]")
for g, a, b in cand[:32]:
    print("#   spheres %d (link %d) - %d (link %d): min gap %.3f" % (a, m.sph_link[a], b, m.sph_link[b], g))
