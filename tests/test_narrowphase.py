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


# ------------------------------------------------------------------------------------------------ GJK / EPA
B2N_POINT, B2N_SEGMENT, B2N_BOX = 0, 1, 2


def _shape(kind, c, R=np.eye(3), h=(0, 0, 0), r=0.0):
    return np.concatenate([[kind], np.asarray(c, np.float64), np.asarray(R, np.float64).reshape(9), np.asarray(h, np.float64),
                           [r]]).astype(np.float32)


def convex_contact(o, A, B, margin=0.01):
    o.lib.b2o_convex_contact.restype = C.c_int
    n = np.zeros(3, np.float32)
    out = np.zeros(8, np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    k = o.lib.b2o_convex_contact(p(A), p(B), C.c_float(margin), p(n), p(out))
    return k, n.astype(np.float64), out.astype(np.float64)


def _point_box_dist(x, c, R, h):
    """Signed distance of point x to a box (negative inside) and the closest surface point."""
    l = np.asarray(R).T @ (x - c)
    cl = np.clip(l, -h, h)
    if np.all(cl == l):
        pen = h - np.abs(l)
        ax = int(np.argmin(pen))
        cl = l.copy()
        cl[ax] = np.sign(l[ax] if l[ax] != 0 else 1.0) * h[ax]
        return -pen[ax], c + np.asarray(R) @ cl
    return np.linalg.norm(l - cl), c + np.asarray(R) @ cl


def _points_box_dist(X, R, h):
    """Signed distances of the points X (rows) to a box centred at the origin (vectorised _point_box_dist)."""
    l = X @ np.asarray(R)
    cl = np.clip(l, -h, h)
    inside = np.all(cl == l, axis=1)
    return np.where(inside, -(h - np.abs(l)).min(axis=1), np.linalg.norm(l - cl, axis=1))


def test_gjk_sphere_box_matches_the_closed_form(lib):
    """Sphere (point core + radius) against a rotated box: distance, normal and witness points equal the closest-point
    closed form used by the sphere families of collide(), both apart (GJK) and with the centre inside the box (EPA)."""
    rng = np.random.RandomState(1)
    n_epa = 0
    for it in range(400):
        h = rng.uniform(0.01, 0.2, 3)
        R = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
        r = rng.uniform(0.005, 0.06)
        l = rng.uniform(-1, 1, 3) * h * (1.6 if it % 4 else 0.95)      # every fourth centre inside the box
        x = R @ l
        sd, foot = _point_box_dist(x, np.zeros(3), R, h)
        k, n, out = convex_contact(lib, _shape(B2N_POINT, x, r=r), _shape(B2N_BOX, np.zeros(3), R, h), margin=10.0)
        assert k == 1
        assert abs(out[6] - (sd - r)) < 2e-5, (it, out[6], sd - r)
        n_epa += int(out[7]) == 1
        if abs(sd) > 1e-3:
            if sd < 0:
                # inside: the least-penetration face may be tied within rounding; check the depth and that n is a face normal
                assert np.abs(np.abs(R.T @ n)).max() > 1 - 1e-4
            else:
                np.testing.assert_allclose(n, (x - foot) / sd, atol=2e-4)
                np.testing.assert_allclose(out[3:6], foot, atol=2e-5)
            np.testing.assert_allclose(out[0:3], x - n * r, atol=2e-5)
    assert n_epa > 50


def test_gjk_capsule_box_against_brute_force(lib):
    """Capsule (segment core + radius) against a box: the distance equals the minimum over the segment of the closed-form
    point-box distance (dense sampling + local refinement), the witness points lie on the two surfaces, and n is the
    direction between them."""
    rng = np.random.RandomState(2)
    n_hit = n_deep = 0
    for it in range(300):
        h = rng.uniform(0.02, 0.3, 3) if it % 2 else np.array([0.025] * 3)
        R = _rot(rng.normal(size=3), rng.uniform(0, np.pi)) if it % 3 else np.eye(3)
        Rs = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
        hl, r = rng.uniform(0.03, 0.15), rng.uniform(0.02, 0.06)
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        # centre placed so that the capsule is between 3 cm of overlap and 3 cm of gap
        ts = np.linspace(-hl, hl, 801)

        def surf_dist(c):
            return _points_box_dist(c[None] + ts[:, None] * Rs[:, 2][None], R, h).min() - r
        lo, hi = 0.0, 1.0
        while surf_dist(u * hi) < 0.03:
            hi *= 1.5
        target = rng.uniform(-0.03, 0.03) if it % 4 else -(r + rng.uniform(0.002, 0.02))   # every fourth: the segment enters
        for _ in range(40):
            mid = 0.5 * (lo + hi)
            if surf_dist(u * mid) < target:
                lo = mid
            else:
                hi = mid
        c = u * hi
        ref = surf_dist(c)
        k, n, out = convex_contact(lib, _shape(B2N_SEGMENT, c, Rs, (0, 0, hl), r), _shape(B2N_BOX, np.zeros(3), R, h), margin=0.05)
        assert k == 1, (it, ref)
        n_hit += 1
        core = ref + r
        if core > 1e-3:        # cores apart: GJK is exact
            assert int(out[7]) == 0
            assert abs(out[6] - ref) < 5e-5, (it, out[6], ref)
            pa, pb = out[0:3], out[3:6]
            np.testing.assert_allclose((pa - pb) @ n, out[6], atol=2e-5)
            assert abs(_point_box_dist(pb, np.zeros(3), R, h)[0]) < 3e-5          # on the box
            t = (pa + n * r - c) @ Rs[:, 2]
            assert abs(t) <= hl + 1e-4 and np.linalg.norm(pa + n * r - c - t * Rs[:, 2]) < 3e-5   # r away from the segment
        elif core < -1e-3:     # the segment itself enters the box: EPA depth = least translation that frees the segment
            n_deep += 1
            assert int(out[7]) == 1
            depth = -(out[6] + r)
            # separating translation along n: afterwards the segment no longer penetrates (brute force), and no sampled
            # direction does it with a clearly shorter translation
            moved = c + n * (depth + 1e-4)
            assert _points_box_dist(moved[None] + ts[:, None] * Rs[:, 2][None], R, h).min() > -2e-4, it
            assert depth <= -core * 3 + 0.02
    assert n_hit == 300 and n_deep > 20, (n_hit, n_deep)


def _sat_depth(cA, RA, hA, cB, RB, hB):
    """Largest separation over the 15 SAT axes (negative = overlap depth): exact for two boxes."""
    axes = [RA[:, i] for i in range(3)] + [RB[:, j] for j in range(3)]
    axes += [np.cross(RA[:, i], RB[:, j]) for i in range(3) for j in range(3)]
    best = -1e9
    for a in axes:
        l = np.linalg.norm(a)
        if l < 1e-6:
            continue
        a = a / l
        best = max(best, max(_support_gap(cA, RA, hA, cB, RB, hB, s * a) for s in (-1, 1)))
    return best


def test_epa_depth_of_overlapping_boxes_equals_the_sat_depth(lib):
    """Two sharp boxes through GJK / EPA (no rounding): for overlapping pairs the EPA depth is the minimum translation
    distance, which for boxes is the best of the 15 separating axes (an independent, closed-form answer); for separated pairs
    the GJK distance is bounded below by the best axis separation and the witness points realise it."""
    rng = np.random.RandomState(3)
    n_over = n_sep = 0
    for it in range(400):
        hA, hB = rng.uniform(0.01, 0.05, 3), rng.uniform(0.01, 0.2, 3)
        RA = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
        RB = _rot(rng.normal(size=3), rng.uniform(0, np.pi)) if it % 3 else np.eye(3)
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        g0 = _support_gap(np.zeros(3), RA, hA, np.zeros(3), RB, hB, u)
        cA = u * (-g0 + (rng.uniform(-0.06, -0.005) if it % 2 else rng.uniform(0.0, 0.03)))
        sat = _sat_depth(cA, RA, hA, np.zeros(3), RB, hB)
        k, n, out = convex_contact(lib, _shape(B2N_BOX, cA, RA, hA), _shape(B2N_BOX, np.zeros(3), RB, hB), margin=1.0)
        assert k == 1
        if sat < -1e-3:
            n_over += 1
            assert int(out[7]) == 1
            assert abs(out[6] - sat) < 1e-4 + 0.01 * abs(sat), (it, out[6], sat)
            gap_n = _support_gap(cA, RA, hA, np.zeros(3), RB, hB, n)
            assert abs(gap_n - sat) < 1e-4 + 0.01 * abs(sat), (it, gap_n, sat)      # n IS a least-penetration direction
        elif sat > 1e-3:
            n_sep += 1
            assert int(out[7]) == 0
            assert out[6] >= sat - 2e-5, (it, out[6], sat)
            gap_n = _support_gap(cA, RA, hA, np.zeros(3), RB, hB, n)
            assert abs(gap_n - out[6]) < 5e-5, (it, gap_n, out[6])                  # the distance is realised along n
            la = np.abs(RA.T @ (out[0:3] - cA)) - hA
            lb = np.abs(RB.T @ out[3:6]) - hB
            assert abs(la.max()) < 3e-5 and abs(lb.max()) < 3e-5                    # witnesses on the two surfaces
    assert n_over > 60 and n_sep > 100, (n_over, n_sep)


def test_gjk_degenerate_configurations(lib):
    """Symmetric / degenerate inputs: concentric shapes, a segment through the box centre along an axis, touching cores."""
    box = _shape(B2N_BOX, [0, 0, 0], np.eye(3), [0.1, 0.2, 0.3])
    k, n, out = convex_contact(lib, _shape(B2N_POINT, [0, 0, 0], r=0.01), box, margin=1.0)
    assert k == 1 and abs(out[6] - (-0.1 - 0.01)) < 1e-5 and abs(abs(n[0]) - 1) < 1e-5
    k, n, out = convex_contact(lib, _shape(B2N_SEGMENT, [0, 0, 0], np.eye(3), (0, 0, 0.5), 0.02), box, margin=1.0)
    assert k == 1 and abs(out[6] - (-0.1 - 0.02)) < 1e-4 and abs(n[2]) < 1e-4 and np.isfinite(out).all()
    k, n, out = convex_contact(lib, _shape(B2N_POINT, [0.1, 0, 0], r=0.01), box, margin=1.0)       # centre exactly on a face
    assert k == 1 and abs(out[6] + 0.01) < 1e-5 and np.isfinite(n).all()
    k, n, out = convex_contact(lib, _shape(B2N_POINT, [0.5, 0, 0], r=0.01), box, margin=0.01)       # far: no contact
    assert k == 0


def test_narrowphase_symmetries(lib):
    """Size-independent properties: swapping the two shapes negates the normal and keeps the depth (box-box and GJK / EPA);
    a rigid motion of both shapes moves the contact with it."""
    rng = np.random.RandomState(7)
    n_bb = n_cc = 0
    for it in range(200):
        hA, hB = rng.uniform(0.01, 0.05, 3), rng.uniform(0.01, 0.2, 3)
        RA = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
        RB = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
        u = rng.normal(size=3)
        u /= np.linalg.norm(u)
        g0 = _support_gap(np.zeros(3), RA, hA, np.zeros(3), RB, hB, u)
        cA = u * (-g0 + rng.uniform(-0.01, 0.006))
        # rigid motion (rotation Q, translation t)
        Q = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
        t = rng.uniform(-0.3, 0.3, 3)
        k1, n1, p1 = box_box(lib, cA, RA, hA, np.zeros(3), RB, hB)
        k2, n2, p2 = box_box(lib, np.zeros(3), RB, hB, cA, RA, hA)
        k3, n3, p3 = box_box(lib, Q @ cA + t, Q @ RA, hA, t, Q @ RB, hB)
        if k1 and k2:
            n_bb += 1
            assert abs(p1[:, 6].min() - p2[:, 6].min()) < 2e-4, (it, p1[:, 6], p2[:, 6])      # same deepest penetration
            assert np.dot(n1, n2) < -0.7 or abs(p1[:, 6].min()) < 2e-3, (it, n1, n2)       # opposite normals (face / edge choice may differ near ties)
        if k1 and k3 and abs(p1[:, 6].min()) > 1e-4:
            assert abs(p1[:, 6].min() - p3[:, 6].min()) < 1e-4, (it, p1[:, 6], p3[:, 6])
            assert np.linalg.norm(Q @ n1.astype(np.float64) - n3) < 2e-2 or k1 != k3, (it, n1, n3)
        # GJK / EPA: capsule vs box, swapped and moved
        Rs = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
        cap = _shape(B2N_SEGMENT, cA, Rs, (0, 0, 0.08), 0.03)
        box = _shape(B2N_BOX, np.zeros(3), RB, hB)
        ka, na, oa = convex_contact(lib, cap, box, margin=0.05)
        kb, nb, ob = convex_contact(lib, box, cap, margin=0.05)
        kc, nc, oc = convex_contact(lib, _shape(B2N_SEGMENT, Q @ cA + t, Q @ Rs, (0, 0, 0.08), 0.03), _shape(B2N_BOX, t, Q @ RB, hB), margin=0.05)
        assert ka == kb == kc
        if ka:
            n_cc += 1
            assert abs(oa[6] - ob[6]) < 1e-4 and abs(oa[6] - oc[6]) < 1e-4, (it, oa[6], ob[6], oc[6])
            if abs(oa[6] + 0.03) > 2e-3:      # away from the cores just touching (normal undefined there)
                assert np.linalg.norm(na + nb) < 2e-2, (it, na, nb)
                assert np.linalg.norm(Q @ na - nc) < 2e-2, (it, na, nc)
                np.testing.assert_allclose(Q @ oa[0:3] + t, oc[0:3], atol=2e-3)
    assert n_bb >= 15 and n_cc > 60, (n_bb, n_cc)
