"""Known-answer and brute-force tests of include/b2env_narrowphase.h (box-box SAT + face clipping) and of the contact
generation built on it (oracle/b2oracle.c collide()).  The CUDA kernel and the oracle SHARE that header, so GPU-vs-oracle
parity says nothing about its correctness — these tests do: closed-form manifolds (face-down cube, cube overhanging the
table rim, pad face on cube face, crossed edges), and separation / penetration depths against a brute-force sampling of the
two boxes.  CPU only.
"""
import ctypes as C

import numpy as np
import pytest

from common import TASK_PUSH, panda_task_setup


def _rot(axis, ang):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    x, y, z = axis
    c, s = np.cos(ang), np.sin(ang)
    t = 1 - c
    return np.array([[t * x * x + c, t * x * y - s * z, t * x * z + s * y],
                     [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
                     [t * x * z - s * y, t * y * z + s * x, t * z * z + c]])


@pytest.fixture(scope="module")
def lib(oracle_lib):
    m, p = panda_task_setup(TASK_PUSH)
    o = oracle_lib.Oracle(m, p, 1)
    o.lib.b2o_box_box.restype = C.c_int
    o.lib.b2o_collide.restype = C.c_int
    return o


def box_box(o, cA, RA, hA, cB, RB, hB, margin=0.01):
    f = lambda x: np.ascontiguousarray(x, np.float32)
    cA, RA, hA, cB, RB, hB = f(cA), f(np.asarray(RA).reshape(9)), f(hA), f(cB), f(np.asarray(RB).reshape(9)), f(hB)
    n = np.zeros(3, np.float32)
    out = np.zeros((4, 8), np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    k = o.lib.b2o_box_box(p(cA), p(RA), p(hA), p(cB), p(RB), p(hB), C.c_float(margin), p(n), p(out))
    return k, n, out[:k]


def test_cube_resting_on_slab_is_the_vertex_face_manifold(lib):
    """Face-down cube well inside a big slab: four points = the cube's bottom vertices, normal +z, depth = gap."""
    a = 0.025
    k, n, pts = box_box(lib, [0.3, 0.1, 0.625 + a - 0.001], np.eye(3), [a] * 3, [0.85, 0, 0.6], np.eye(3), [0.75, 0.5, 0.025])
    assert k == 4
    np.testing.assert_allclose(n, [0, 0, 1], atol=1e-6)           # from the slab (B) towards the cube (A)
    np.testing.assert_allclose(pts[:, 6], -0.001, atol=1e-6)
    got = sorted((round(float(x), 4), round(float(y), 4)) for x, y in pts[:, :2])
    assert got == sorted((round(0.3 + sx * a, 4), round(0.1 + sy * a, 4)) for sx in (-1, 1) for sy in (-1, 1))
    np.testing.assert_allclose(pts[:, 2], 0.625 - 0.001, atol=1e-6)  # on the cube
    np.testing.assert_allclose(pts[:, 5], 0.625, atol=1e-6)          # foot on the slab top
    assert len(set(pts[:, 7])) == 4


def test_cube_overhanging_the_rim(lib):
    """Cube pushed 1 cm past the table edge x = 0.1: the bottom face is clipped by the rim line — two cube vertices inside
    plus the two intersections of the rim with the cube's bottom edges (x = 0.1)."""
    a = 0.025
    k, n, pts = box_box(lib, [0.1 + 0.015, 0.0, 0.625 + a], np.eye(3), [a] * 3, [0.85, 0, 0.6], np.eye(3), [0.75, 0.5, 0.025])
    assert k == 4
    np.testing.assert_allclose(n, [0, 0, 1], atol=1e-6)
    xs = sorted(round(float(x), 5) for x in pts[:, 0])
    assert xs == [0.1, 0.1, 0.14, 0.14]
    np.testing.assert_allclose(np.abs(pts[:, 1]), a, atol=1e-6)
    np.testing.assert_allclose(pts[:, 6], 0.0, atol=1e-6)
    # two kinds of points, two feature classes: corners of the smaller face (incident vertices 0..3 or reference corners 20..23,
    # depending on which box owns the reference face — a tie here, A wins) and edge x clip-line crossings (4..19)
    labels = sorted(int(i) % 32 for i in pts[:, 7])
    assert sum(l < 4 or 20 <= l < 24 for l in labels) == 2 and sum(4 <= l < 20 for l in labels) == 2, labels
    assert len(set(labels)) == 4
    # tilted over the rim (rotation about y through the rim line): still a face contact of the table top, now 2 points deep
    R = _rot([0, 1, 0], 0.2)
    c = np.array([0.1, 0, 0.625]) + R @ np.array([0.01, 0, a])
    k, n, pts = box_box(lib, c, R, [a] * 3, [0.85, 0, 0.6], np.eye(3), [0.75, 0.5, 0.025])
    assert k >= 1 and n[2] > 0.9


def test_pad_face_on_cube_face_gives_four_points(lib):
    """A finger pad (18 x 16 x 54 mm box) pressed flat against a cube side face: four-point manifold whose points are the
    pad face's vertices that lie inside the cube face, normal along the face axis."""
    a = 0.025
    hp = [0.009, 0.008, 0.027]
    # pad centred on the cube's +y face, its own -y face touching with 0.5 mm overlap; pad long axis (z) vertical
    cpad = [0.0, a + hp[1] - 0.0005, 0.0]
    k, n, pts = box_box(lib, cpad, np.eye(3), hp, [0, 0, 0], np.eye(3), [a] * 3)
    assert k == 4
    np.testing.assert_allclose(n, [0, 1, 0], atol=1e-6)       # from the cube towards the pad
    np.testing.assert_allclose(pts[:, 6], -0.0005, atol=1e-6)
    # the pad face is 18 x 54 mm, the cube face 50 x 50: clipped to |z| <= 25 mm, |x| <= 9 mm
    assert sorted(round(abs(float(x)), 4) for x in pts[:, 0]) == [0.009] * 4
    assert sorted(round(abs(float(z)), 4) for z in pts[:, 2]) == [0.025] * 4


def test_crossed_edges_give_one_point(lib):
    """Two boxes touching edge to edge (each rotated 45 deg about a different axis): one contact, normal along the common
    perpendicular, points on the two edges."""
    a = 0.025
    RA = _rot([1, 0, 0], np.pi / 4)     # edge along x on top
    RB = _rot([0, 1, 0], np.pi / 4)     # edge along y at the bottom
    d = 2 * a * np.sqrt(2) - 0.001      # centre distance along z: 1 mm overlap
    k, n, pts = box_box(lib, [0, 0, d], RB, [a] * 3, [0, 0, 0], RA, [a] * 3)
    assert k == 1
    np.testing.assert_allclose(np.abs(n), [0, 0, 1], atol=1e-5)
    assert n[2] > 0                                   # from B (below) towards A (above)
    np.testing.assert_allclose(pts[0, 6], -0.001, atol=2e-6)
    np.testing.assert_allclose(pts[0, :2], [0, 0], atol=1e-5)
    assert int(pts[0, 7]) >= 384


def _support_gap(cA, RA, hA, cB, RB, hB, n):
    """Separation of the two boxes along direction n (unit, from B to A): min over A of x.n minus max over B of x.n."""
    sg = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float64)
    va = cA + (sg * hA) @ np.asarray(RA).T
    vb = cB + (sg * hB) @ np.asarray(RB).T
    return (va @ n).min() - (vb @ n).max()


def test_random_pairs_against_brute_force(lib):
    """2000 random box pairs near contact: (1) no contact is reported iff some axis separates by more than the margin (checked
    against a dense direction sampling); (2) the reported normal is a valid separating / least-penetration direction: the
    support gap along it equals the deepest reported distance; (3) every reported point pair lies on the two boxes and
    pa - pb is parallel to the normal."""
    rng = np.random.RandomState(0)
    dirs = rng.normal(size=(4000, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    n_hit = n_miss = n_spec_missed = 0
    for it in range(2000):
        hA = rng.uniform(0.008, 0.03, 3)
        hB = rng.uniform(0.02, 0.3, 3) if it % 2 else rng.uniform(0.008, 0.03, 3)
        RA = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
        RB = _rot(rng.normal(size=3), rng.uniform(0, np.pi)) if it % 3 else np.eye(3)
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        # place A so that the gap along u is in [-6 mm, +8 mm]
        g0 = _support_gap(np.zeros(3), RA, hA, np.zeros(3), RB, hB, u)
        cA = u * (-g0 + rng.uniform(-0.006, 0.008))
        k, n, pts = box_box(lib, cA, RA, hA, np.zeros(3), RB, hB, margin=0.01)
        gaps = np.array([_support_gap(cA, RA, hA, np.zeros(3), RB, hB, d) for d in dirs[:200]])
        if k == 0:
            n_miss += 1
            # truly separated by more than the margin along SOME axis (face normals or edge crosses): brute-force over a dense set
            cand = [RA[:, i] for i in range(3)] + [RB[:, j] for j in range(3)]
            cand += [np.cross(RA[:, i], RB[:, j]) for i in range(3) for j in range(3)]
            best = max(max(_support_gap(cA, RA, hA, np.zeros(3), RB, hB, s * c / np.linalg.norm(c)) for s in (-1, 1))
                       for c in cand if np.linalg.norm(c) > 1e-6)
            # no contact reported: the boxes must not overlap (some axis separates them).  They may still be closer than the
            # margin when the closest features are vertex-edge / vertex-vertex (the clipped incident face is empty): such
            # speculative contacts are not generated, as in Bullet's box-box detector; counted below
            assert best > -1e-6, (it, best)
            n_spec_missed += best <= 0.01 - 1e-5
            continue
        n_hit += 1
        n = n.astype(np.float64)
        assert abs(np.linalg.norm(n) - 1) < 1e-4
        gap_n = _support_gap(cA, RA, hA, np.zeros(3), RB, hB, n)
        assert gap_n <= 0.01 + 1e-5
        # no sampled direction separates better than the chosen one by more than the 5 % edge / face preference + sampling error
        assert gaps.max() <= max(gap_n, 0) + 0.3 * abs(gap_n) + 4e-3, (it, gaps.max(), gap_n)
        deepest = pts[:, 6].min()
        # no point can be deeper than the support gap along the normal.  The deepest vertex may lie beside the reference face
        # and be clipped away (face clipping then reports the depth at the clip line instead): overlapping boxes must still
        # yield a penetrating point, at least half as deep as the support gap
        assert deepest >= gap_n - 2e-5, (it, deepest, gap_n)
        if gap_n < -1e-4:
            assert deepest < 1e-5 and deepest <= 0.5 * gap_n + 1e-4, (it, deepest, gap_n)
        for q in range(k):
            pa, pb, dist = pts[q, :3].astype(np.float64), pts[q, 3:6].astype(np.float64), float(pts[q, 6])
            la = np.abs(RA.T @ (pa - cA)) - hA
            lb = np.abs(RB.T @ pb) - hB
            assert la.max() < 2e-5 and lb.max() < 2e-5, (it, q, la, lb)         # inside / on the boxes
            assert min(abs(la.max()), abs(lb.max())) < 2e-5                     # at least one of them ON a surface
            d = pa - pb
            assert abs(d @ n - dist) < 1e-5, (it, q, d, n, dist)
            if int(pts[q, 7]) % 1024 < 384 or dist < 0:   # face points by construction; edge-edge points unless the closest
                #                                           points were clamped to the ends of separated edges
                assert np.linalg.norm(d - (d @ n) * n) < 2e-4 + 0.5 * abs(dist), (it, q, d, n, dist)
            assert dist < 0.01
    print("box-box brute force: %d pairs with contacts, %d without (%d of them closer than the margin through a vertex-edge / "
          "vertex-vertex feature)" % (n_hit, n_miss, n_spec_missed))
    assert n_hit > 300 and n_miss > 100, (n_hit, n_miss)
    assert n_spec_missed < 0.25 * n_hit, (n_spec_missed, n_hit)


def _collide(o, m, p, q, obj_pose):
    out = np.zeros((12, 16), np.float32)
    ov = C.c_int(0)
    q = np.ascontiguousarray(q, np.float32)
    pose = np.ascontiguousarray(obj_pose, np.float32)
    n = o.lib.b2o_collide(C.byref(m), C.byref(p), q.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.c_void_p),
                          out.ctypes.data_as(C.c_void_p), C.byref(ov))
    return out[:n], ov.value


def test_world_contact_sets(lib):
    """The contact families of oracle collide() on hand-placed configurations: resting cube (fast path, keys 0..7), cube at
    the rim (general box-box vs the top slab), cube leaning against a table leg, cube on the ground plane."""
    m, p = panda_task_setup(TASK_PUSH)
    home = np.array([m.home[i] for i in range(9)], np.float32)
    # resting on the table, far from the rim
    c, ov = _collide(lib, m, p, home, [0.45, 0.0, 0.65, 0, 0, 0, 1])
    assert sorted(c[:, 0].astype(int)) == [0, 1, 2, 3] and not ov
    np.testing.assert_allclose(c[:, 13], 0.0, atol=1e-6)
    # 1 cm over the rim at x = 0.1: general path, four points, keys in the cube-vs-static-box-0 range
    c, ov = _collide(lib, m, p, home, [0.115, 0.0, 0.65, 0, 0, 0, 1])
    keys = c[:, 0].astype(int)
    assert len(keys) == 4 and all(4096 <= k < 4096 + 1024 for k in keys)
    assert sorted(round(float(x), 4) for x in c[:, 4]) == [0.1, 0.1, 0.14, 0.14]
    np.testing.assert_allclose(c[:, 10:13], [[0, 0, 1]] * 4, atol=1e-6)
    # against the leg at (0.2, -0.4): cube on the ground touching the leg's +y face
    c, ov = _collide(lib, m, p, home, [0.2, -0.4 + 0.05 + 0.025 - 0.0005, 0.025, 0, 0, 0, 1])
    keys = c[:, 0].astype(int)
    leg = [k for k in keys if 4096 + 1024 <= k < 4096 + 5 * 1024]
    plane = [k for k in keys if 8 <= k < 16]
    assert len(leg) == 4 and len(plane) == 4, keys
    ln = c[[k in leg for k in keys], 10:13]
    np.testing.assert_allclose(ln, [[0, 1, 0]] * 4, atol=1e-6)       # from the leg towards the cube
    np.testing.assert_allclose(c[[k in leg for k in keys], 13], -0.0005, atol=1e-6)
