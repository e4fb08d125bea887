/*
 * b2env_narrowphase.h — convex narrowphase routines of the step pipeline: box-box (separating axes + face clipping,
 * edge-edge closest points) and GJK / EPA for convex proxies given by support mappings (point, segment, box, each
 * with a rounding radius).
 *
 * What this replaces: the narrowphase p.stepSimulation() runs inside Bullet (btBoxBoxDetector for box pairs,
 * btGjkPairDetector + btGjkEpaPenetrationDepthSolver for convex hulls) for the bodies the reference loads at
 * pybullet_robot_envs/envs/world_envs/world_env.py:62-66 (plane, table) and :81-84 (object), and for the Panda links,
 * loaded with URDF_USE_SELF_COLLISION at envs/panda_envs/panda_env.py:53 [Bullet internals: EXT-recalled, SURVEY.md
 * Appendix D.2].  Written from the published algorithms (separating-axis theorem, Sutherland-Hodgman clipping,
 * Gilbert-Johnson-Keerthi distance, expanding polytope), not from Bullet's sources.
 *
 * ONE scalar implementation, included by BOTH the CUDA kernel (csrc/b2env.cu: one lane runs one candidate pair) and the
 * CPU oracle (oracle/b2oracle.c): discrete decisions (which axis separates least, which clipped points survive) must
 * be identical on both sides for contact keys to match bit for bit.  Because the two sides share this source, the
 * GPU-vs-oracle tests do NOT verify it; it is pinned independently by tests/test_narrowphase.py (closed-form manifolds,
 * brute-force penetration depths, support-sampling checks of GJK / EPA).
 *
 * Conventions: rotation matrices row-major, world <- body; a contact has a point on A, a point on B, the normal n
 * pointing from B towards A, and dist = (pA - pB) . n (negative = penetration); ids are topological (which vertex /
 * edge / clip line produced the point) so that they are stable from step to step (warm start by key).
 *
 * Configure before including:  B2N_REAL (float | double),  B2N_FN (function qualifiers, e.g. static inline or
 * __device__ __noinline__).  C99 / C++ / CUDA.
 */
#ifndef B2ENV_NARROWPHASE_H
#define B2ENV_NARROWPHASE_H

#ifndef B2N_REAL
#define B2N_REAL float
#endif
#ifndef B2N_FN
#define B2N_FN static inline
#endif
#ifndef B2N_SQRT
#define B2N_SQRT(x) ((B2N_REAL)sqrt((double)(x)))
#endif
#ifndef B2N_FABS
#define B2N_FABS(x) ((x) < 0 ? -(x) : (x))
#endif

#define B2N_MAX_POINTS 4   /* manifold points kept per pair (Bullet's persistent manifold size) */
#define B2N_ID_STRIDE 1024 /* ids of one pair.  Face contact: 32 * reference face (0-5 faces of A, 6-11 of B) + label, label =
                              0-3 incident vertex | 4-19 incident edge x clip line | 20-23 reference corner;
                              edge-edge: 384 + 16 * (3 i + j) + edge signs                                          */

typedef struct b2n_contact {
  B2N_REAL pa[3], pb[3]; /* points on A / on B (world)        */
  B2N_REAL dist;         /* (pa - pb) . n, negative = overlap */
  int id;
} b2n_contact;

/* ---------------------------------------------------------------------------------------------- box - box */
/* Candidate axis bookkeeping: s = separation along the axis (negative = overlap depth). */
B2N_FN int b2n_box_box(const B2N_REAL* cA, const B2N_REAL* RA, const B2N_REAL* hA, const B2N_REAL* cB, const B2N_REAL* RB,
                       const B2N_REAL* hB, B2N_REAL margin, B2N_REAL* n_out, b2n_contact* out) {
  typedef B2N_REAL real;
  const real eps = (real)1e-6;
  real t[3] = {cB[0] - cA[0], cB[1] - cA[1], cB[2] - cA[2]};
  real tA[3], R[3][3], Q[3][3];
  int i, j, k;
  for (i = 0; i < 3; i++) tA[i] = RA[i] * t[0] + RA[3 + i] * t[1] + RA[6 + i] * t[2];          /* RA^T t */
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      R[i][j] = RA[i] * RB[j] + RA[3 + i] * RB[3 + j] + RA[6 + i] * RB[6 + j];                /* RA^T RB */
      Q[i][j] = B2N_FABS(R[i][j]) + eps;
    }
  real best = (real)-1e30;
  int code = -1;  /* 0-2: face of A, 3-5: face of B, 6-14: edge i x j */
  real bsign = 1;
  /* faces of A */
  for (i = 0; i < 3; i++) {
    real s = B2N_FABS(tA[i]) - (hA[i] + Q[i][0] * hB[0] + Q[i][1] * hB[1] + Q[i][2] * hB[2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; bsign = tA[i] < 0 ? (real)-1 : (real)1; }
  }
  /* faces of B */
  real tB[3];
  for (j = 0; j < 3; j++) {
    tB[j] = tA[0] * R[0][j] + tA[1] * R[1][j] + tA[2] * R[2][j];
    real s = B2N_FABS(tB[j]) - (hB[j] + Q[0][j] * hA[0] + Q[1][j] * hA[1] + Q[2][j] * hA[2]);
    if (s > margin) return 0;
    if (s > best + (real)1e-5) { best = s; code = 3 + j; bsign = tB[j] < 0 ? (real)-1 : (real)1; }  /* ties: faces of A first */
  }
  /* edge x edge: axis = a_i x b_j (in A's frame: e_i x R[:,j]); a face axis is preferred unless the edge axis is
   * clearly better (5 % + 1e-4), which keeps resting / sliding face contacts on the face path */
  real ebest = (real)-1e30, en[3] = {0, 0, 0};
  int ecode = -1;
  for (i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      real l2 = (real)1 - R[i][j] * R[i][j];
      if (l2 < (real)1e-6) continue;   /* parallel edges: covered by the face axes */
      real l = B2N_SQRT(l2);
      real d = tA[i2] * R[i1][j] - tA[i1] * R[i2][j];
      real ra = hA[i1] * Q[i2][j] + hA[i2] * Q[i1][j];
      real rb = hB[j1] * Q[i][j2] + hB[j2] * Q[i][j1];
      real s = (B2N_FABS(d) - (ra + rb)) / l;
      if (s > margin) return 0;
      if (s > ebest) {
        ebest = s; ecode = 6 + 3 * i + j;
        /* axis in A's frame: e_i x b_j = (0,..) ; components: [i1] = -R[i2][j], [i2] = R[i1][j] */
        real sg = d < 0 ? (real)-1 : (real)1;
        en[i] = 0; en[i1] = -R[i2][j] * sg / l; en[i2] = R[i1][j] * sg / l;
      }
    }
  }
  if (ecode >= 0 && ebest > best + (real)0.05 * B2N_FABS(best) + (real)1e-4) { best = ebest; code = ecode; }

  if (code >= 6) {
    /* ---- edge-edge: one point.  Normal (A -> B) in world = RA * en */
    real nw[3];
    for (k = 0; k < 3; k++) nw[k] = RA[3 * k] * en[0] + RA[3 * k + 1] * en[1] + RA[3 * k + 2] * en[2];
    const int ia = (code - 6) / 3, jb = (code - 6) % 3;
    /* the edge of A furthest along +n, the edge of B furthest along -n */
    real pa[3] = {cA[0], cA[1], cA[2]}, pb[3] = {cB[0], cB[1], cB[2]};
    int sa = 0, sb = 0;
    for (k = 0; k < 3; k++) {
      if (k != ia) {
        real dk = nw[0] * RA[k] + nw[1] * RA[3 + k] + nw[2] * RA[6 + k];
        real sg = dk > 0 ? (real)1 : (real)-1;
        if (sg > 0) sa |= 1 << k;
        pa[0] += sg * hA[k] * RA[k]; pa[1] += sg * hA[k] * RA[3 + k]; pa[2] += sg * hA[k] * RA[6 + k];
      }
      if (k != jb) {
        real dk = nw[0] * RB[k] + nw[1] * RB[3 + k] + nw[2] * RB[6 + k];
        real sg = dk > 0 ? (real)-1 : (real)1;
        if (sg > 0) sb |= 1 << k;
        pb[0] += sg * hB[k] * RB[k]; pb[1] += sg * hB[k] * RB[3 + k]; pb[2] += sg * hB[k] * RB[6 + k];
      }
    }
    /* closest points of the two lines pa + al ua, pb + be ub */
    real ua[3] = {RA[ia], RA[3 + ia], RA[6 + ia]}, ub[3] = {RB[jb], RB[3 + jb], RB[6 + jb]};
    real w[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    real uaub = ua[0] * ub[0] + ua[1] * ub[1] + ua[2] * ub[2];
    real q1 = ua[0] * w[0] + ua[1] * w[1] + ua[2] * w[2], q2 = -(ub[0] * w[0] + ub[1] * w[1] + ub[2] * w[2]);
    real dd = (real)1 - uaub * uaub;
    real al = 0, be = 0;
    if (dd > (real)1e-6) { al = (q1 + uaub * q2) / dd; be = (uaub * q1 + q2) / dd; }
    al = al > hA[ia] ? hA[ia] : (al < -hA[ia] ? -hA[ia] : al);
    be = be > hB[jb] ? hB[jb] : (be < -hB[jb] ? -hB[jb] : be);
    for (k = 0; k < 3; k++) { out[0].pa[k] = pa[k] + al * ua[k]; out[0].pb[k] = pb[k] + be * ub[k]; }
    n_out[0] = -nw[0]; n_out[1] = -nw[1]; n_out[2] = -nw[2];   /* from B towards A */
    out[0].dist = (out[0].pa[0] - out[0].pb[0]) * n_out[0] + (out[0].pa[1] - out[0].pb[1]) * n_out[1] +
                  (out[0].pa[2] - out[0].pb[2]) * n_out[2];
    out[0].id = 384 + (code - 6) * 16 + ((sa >> (ia == 0 ? 1 : 0)) & 1) * 8 + ((sa >> (ia == 2 ? 1 : 2)) & 1) * 4 +
                ((sb >> (jb == 0 ? 1 : 0)) & 1) * 2 + ((sb >> (jb == 2 ? 1 : 2)) & 1);
    return 1;
  }

  /* ---- face contact: reference box X (the one owning the axis), incident box Y */
  const int refA = code < 3;
  const int kx = refA ? code : code - 3;
  const real* cX = refA ? cA : cB; const real* RX = refA ? RA : RB; const real* hX = refA ? hA : hB;
  const real* cY = refA ? cB : cA; const real* RY = refA ? RB : RA; const real* hY = refA ? hB : hA;
  /* reference normal (world), pointing from X towards Y */
  real sgn = refA ? bsign : -bsign;
  real nr[3] = {sgn * RX[kx], sgn * RX[3 + kx], sgn * RX[6 + kx]};
  /* incident face of Y: most anti-parallel to nr */
  int jy = 0;
  real dmax = -1, dj[3];
  for (j = 0; j < 3; j++) {
    dj[j] = nr[0] * RY[j] + nr[1] * RY[3 + j] + nr[2] * RY[6 + j];
    if (B2N_FABS(dj[j]) > dmax + (real)1e-6) { dmax = B2N_FABS(dj[j]); jy = j; }
  }
  const real ty = dj[jy] > 0 ? (real)-1 : (real)1;
  const int j1 = (jy + 1) % 3, j2 = (jy + 2) % 3, k1 = (kx + 1) % 3, k2 = (kx + 2) % 3;
  /* incident face vertices v = 0..3 (signs (-,-),(+,-),(+,+),(-,+) along j1, j2), in X's face frame: (x, y) along X's
   * axes k1, k2 and h = height above the reference face */
  real px[8], py[8], ph[8];
  int lab[8], elab[8];   /* vertex label; label of the edge from this vertex to the next (0-3 incident edge, 4-7 clip line) */
  int np = 4;
  {
    real fc[3];
    for (k = 0; k < 3; k++) fc[k] = cY[k] + ty * hY[jy] * RY[3 * k + jy] - cX[k];
    const real sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
    for (i = 0; i < 4; i++) {
      real v[3];
      for (k = 0; k < 3; k++) v[k] = fc[k] + sx[i] * hY[j1] * RY[3 * k + j1] + sy[i] * hY[j2] * RY[3 * k + j2];
      px[i] = v[0] * RX[k1] + v[1] * RX[3 + k1] + v[2] * RX[6 + k1];
      py[i] = v[0] * RX[k2] + v[1] * RX[3 + k2] + v[2] * RX[6 + k2];
      ph[i] = v[0] * nr[0] + v[1] * nr[1] + v[2] * nr[2] - hX[kx];
      lab[i] = i; elab[i] = i;
    }
  }
  /* Sutherland-Hodgman against the four sides of the reference face; clip line c = 0..3: x <= hx, x >= -hx, y <= hy, y >= -hy */
  for (int c = 0; c < 4; c++) {
    real qx[8], qy[8], qh[8];
    int ql[8], qe[8], nq = 0;
    const real lim = (c < 2) ? hX[k1] : hX[k2];
    for (i = 0; i < np; i++) {
      const int i2 = (i + 1 == np) ? 0 : i + 1;
      const real a0 = ((c < 2) ? px[i] : py[i]) * ((c & 1) ? (real)-1 : (real)1) - lim;    /* <= 0: inside */
      const real a1 = ((c < 2) ? px[i2] : py[i2]) * ((c & 1) ? (real)-1 : (real)1) - lim;
      const int in0 = a0 <= 0, in1 = a1 <= 0;
      if (in0 && nq < 8) { qx[nq] = px[i]; qy[nq] = py[i]; qh[nq] = ph[i]; ql[nq] = lab[i]; qe[nq] = elab[i]; nq++; }
      if (in0 != in1 && nq < 8) {
        const real f = a0 / (a0 - a1);
        qx[nq] = px[i] + f * (px[i2] - px[i]); qy[nq] = py[i] + f * (py[i2] - py[i]); qh[nq] = ph[i] + f * (ph[i2] - ph[i]);
        /* crossing of subject edge `elab[i]` with clip line c: incident edge e -> 4 + 4 e + c; an earlier clip line c' ->
         * reference corner (c', c) -> 20 + corner index */
        if (elab[i] < 4) ql[nq] = 4 + 4 * elab[i] + c;
        else { const int c0 = elab[i] - 4; ql[nq] = 20 + ((c0 & 1) | ((c & 1) << 1)); }
        /* leaving the window: the next output edge runs along clip line c; entering: it continues on the subject edge */
        qe[nq] = in0 ? 4 + c : elab[i];
        nq++;
      }
    }
    /* an inside vertex followed by an outside one: the edge from the inside vertex still lies on its subject edge up to the
     * crossing (label already copied above) */
    np = nq;
    for (i = 0; i < np; i++) { px[i] = qx[i]; py[i] = qy[i]; ph[i] = qh[i]; lab[i] = ql[i]; elab[i] = qe[i]; }
    if (np == 0) return 0;
  }
  /* keep the points within the margin of the reference face */
  int m = 0;
  for (i = 0; i < np; i++)
    if (ph[i] < margin) { px[m] = px[i]; py[m] = py[i]; ph[m] = ph[i]; lab[m] = lab[i]; m++; }
  if (m == 0) return 0;
  /* reduce to four: deepest, farthest from it, farthest from that line on either side (ties within 1e-5: lowest label) */
  int keep[4], nk = 0;
  if (m <= B2N_MAX_POINTS) {
    for (i = 0; i < m; i++) keep[nk++] = i;
  } else {
    /* ties (symmetric configurations: a square face seen from one of its corners) are decided by the lowest label; "equal"
     * means within 1e-5 m for depths and within 1e-4 relative for squared distances / areas, far above rounding, so that
     * both builds of this header (CPU oracle, CUDA with contracted FMAs) take the same decisions                       */
    const real tol = (real)1e-5, rtol = (real)1e-4, atol = (real)1e-12;
    int b0 = 0;
    for (i = 1; i < m; i++)
      if (ph[i] < ph[b0] - tol || (B2N_FABS(ph[i] - ph[b0]) <= tol && lab[i] < lab[b0])) b0 = i;
    int b1 = -1; real v1 = -1;
    for (i = 0; i < m; i++) {
      if (i == b0) continue;
      real d2 = (px[i] - px[b0]) * (px[i] - px[b0]) + (py[i] - py[b0]) * (py[i] - py[b0]);
      const real eq = rtol * (d2 > v1 ? d2 : v1) + atol;
      if (b1 < 0 || d2 > v1 + eq || (B2N_FABS(d2 - v1) <= eq && lab[i] < lab[b1])) { b1 = i; v1 = d2; }
    }
    const real ex = px[b1] - px[b0], ey = py[b1] - py[b0];
    int b2 = -1, b3 = -1; real v2 = tol * B2N_SQRT(v1 > 0 ? v1 : 0), v3 = v2;
    for (i = 0; i < m; i++) {
      if (i == b0 || i == b1) continue;
      real cr = ex * (py[i] - py[b0]) - ey * (px[i] - px[b0]);
      if (cr > 0) {
        const real eq = rtol * (cr > v2 ? cr : v2) + atol;
        if (b2 < 0 ? cr > v2 : (cr > v2 + eq || (B2N_FABS(cr - v2) <= eq && lab[i] < lab[b2]))) { b2 = i; v2 = cr; }
      } else {
        const real eq = rtol * (-cr > v3 ? -cr : v3) + atol;
        if (b3 < 0 ? -cr > v3 : (-cr > v3 + eq || (B2N_FABS(-cr - v3) <= eq && lab[i] < lab[b3]))) { b3 = i; v3 = -cr; }
      }
    }
    keep[nk++] = b0; keep[nk++] = b1;
    if (b2 >= 0) keep[nk++] = b2;
    if (b3 >= 0) keep[nk++] = b3;
  }
  /* emit in label order (canonical) */
  for (i = 0; i < nk; i++)
    for (j = i + 1; j < nk; j++)
      if (lab[keep[j]] < lab[keep[i]]) { const int tmp = keep[i]; keep[i] = keep[j]; keep[j] = tmp; }
  for (i = 0; i < nk; i++) {
    const int q = keep[i];
    real pY[3], pXs[3];
    for (k = 0; k < 3; k++) {
      pY[k] = cX[k] + px[q] * RX[3 * k + k1] + py[q] * RX[3 * k + k2] + (hX[kx] + ph[q]) * nr[k];   /* on the incident face */
      pXs[k] = pY[k] - ph[q] * nr[k];                                                                /* its foot on the reference face */
    }
    for (k = 0; k < 3; k++) { out[i].pa[k] = refA ? pXs[k] : pY[k]; out[i].pb[k] = refA ? pY[k] : pXs[k]; }
    out[i].dist = ph[q];
    out[i].id = 32 * ((refA ? 0 : 6) + kx * 2 + (sgn > 0 ? 0 : 1)) + lab[q];   /* a change of reference face starts fresh impulses */
  }
  /* n from B towards A: nr points from X to Y */
  n_out[0] = refA ? -nr[0] : nr[0]; n_out[1] = refA ? -nr[1] : nr[1]; n_out[2] = refA ? -nr[2] : nr[2];
  return nk;
}

/* ---------------------------------------------------------------------------------------------- GJK / EPA */
/* Convex shapes given by a support mapping of a CORE (point, segment, box) plus a rounding radius r: sphere = point + r,
 * capsule = segment + r, rounded box = box + r.  GJK finds the closest points of the two cores; while the cores are apart
 * the contact of the rounded shapes follows directly (this is also how Bullet treats collision margins).  When the cores
 * themselves overlap, EPA expands a polytope inside the Minkowski difference to the nearest boundary face: penetration
 * depth, direction and witness points.  One contact point per pair.                                                       */
#define B2N_POINT 0
#define B2N_SEGMENT 1 /* endpoints c -+ h[2] * (column 2 of R) */
#define B2N_BOX 2
typedef struct b2n_shape {
  int type;
  B2N_REAL c[3]; /* centre (world)                                                            */
  B2N_REAL R[9]; /* world <- body, row-major (box axes = columns; segment axis = column 2)    */
  B2N_REAL h[3]; /* box: half extents; segment: h[2] = half length                            */
  B2N_REAL r;    /* rounding radius                                                           */
} b2n_shape;

#define B2N_GJK_ITERS 32
#define B2N_EPA_ITERS 24
#define B2N_EPA_MAXV (4 + B2N_EPA_ITERS)
#define B2N_EPA_MAXF (4 + 2 * B2N_EPA_ITERS)
#define B2N_EPA_MAXE 32

B2N_FN void b2n_support(const b2n_shape* s, const B2N_REAL* d, B2N_REAL* out) {
  int k, j;
  out[0] = s->c[0]; out[1] = s->c[1]; out[2] = s->c[2];
  if (s->type == B2N_POINT) return;
  for (k = (s->type == B2N_SEGMENT ? 2 : 0); k < 3; k++) {
    const B2N_REAL dk = d[0] * s->R[k] + d[1] * s->R[3 + k] + d[2] * s->R[6 + k];
    const B2N_REAL e = dk >= 0 ? s->h[k] : -s->h[k];
    for (j = 0; j < 3; j++) out[j] += e * s->R[3 * j + k];
  }
}

#define B2N_DOT(a, b) ((a)[0] * (b)[0] + (a)[1] * (b)[1] + (a)[2] * (b)[2])
#define B2N_CROSS(o, a, b)                                                                                             \
  {                                                                                                                    \
    (o)[0] = (a)[1] * (b)[2] - (a)[2] * (b)[1];                                                                       \
    (o)[1] = (a)[2] * (b)[0] - (a)[0] * (b)[2];                                                                       \
    (o)[2] = (a)[0] * (b)[1] - (a)[1] * (b)[0];                                                                       \
  }

/* closest point to the origin on the triangle (a, b, c): barycentric coordinates l[3] (Voronoi-region tests) */
B2N_FN void b2n_closest_triangle(const B2N_REAL* a, const B2N_REAL* b, const B2N_REAL* c, B2N_REAL* l) {
  typedef B2N_REAL real;
  real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  const real d1 = -B2N_DOT(ab, a), d2 = -B2N_DOT(ac, a);
  l[0] = l[1] = l[2] = 0;
  if (d1 <= 0 && d2 <= 0) { l[0] = 1; return; }
  const real d3 = -B2N_DOT(ab, b), d4 = -B2N_DOT(ac, b);
  if (d3 >= 0 && d4 <= d3) { l[1] = 1; return; }
  const real vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { const real v = d1 / (d1 - d3); l[0] = 1 - v; l[1] = v; return; }
  const real d5 = -B2N_DOT(ab, c), d6 = -B2N_DOT(ac, c);
  if (d6 >= 0 && d5 <= d6) { l[2] = 1; return; }
  const real vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { const real w = d2 / (d2 - d6); l[0] = 1 - w; l[2] = w; return; }
  const real va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { const real w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); l[1] = 1 - w; l[2] = w; return; }
  {
    const real den = 1 / (va + vb + vc), v = vb * den, w = vc * den;
    l[0] = 1 - v - w; l[1] = v; l[2] = w;
  }
}

/* Closest point of the simplex W[0..n-1] (n <= 4) to the origin: barycentric coordinates l[]; returns 1 if the origin is
 * inside the tetrahedron.                                                                                              */
B2N_FN int b2n_closest_simplex(B2N_REAL W[4][3], int n, B2N_REAL* l) {
  typedef B2N_REAL real;
  int i, f;
  l[0] = l[1] = l[2] = l[3] = 0;
  if (n == 1) { l[0] = 1; return 0; }
  if (n == 2) {
    real ab[3] = {W[1][0] - W[0][0], W[1][1] - W[0][1], W[1][2] - W[0][2]};
    const real den = B2N_DOT(ab, ab);
    real t = den > 0 ? -B2N_DOT(W[0], ab) / den : 0;
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    l[0] = 1 - t; l[1] = t;
    return 0;
  }
  if (n == 3) { b2n_closest_triangle(W[0], W[1], W[2], l); return 0; }
  {
    /* tetrahedron: the origin is inside if, for every face, it lies on the same side of the face plane as the 4th vertex
     * (strictly, and the tetrahedron is not flat); otherwise the closest point is the nearest of the closest points of the
     * four faces (all four are evaluated: sign tests of a nearly flat tetrahedron are not reliable)                      */
    const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
    real best = (real)1e30;
    int inside = 1;
    for (f = 0; f < 4; f++) {
      const real* a = W[F[f][0]]; const real* b = W[F[f][1]]; const real* c = W[F[f][2]]; const real* d = W[F[f][3]];
      real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, nn[3];
      B2N_CROSS(nn, ab, ac);
      real ad[3] = {d[0] - a[0], d[1] - a[1], d[2] - a[2]};
      const real sp = -B2N_DOT(a, nn), sd = B2N_DOT(ad, nn);
      const real flat = (real)1e-6 * B2N_SQRT(B2N_DOT(nn, nn) * B2N_DOT(ad, ad));
      if (!(sp * sd > 0) || B2N_FABS(sd) <= flat) inside = 0;
    }
    if (inside) return 1;
    for (f = 0; f < 4; f++) {
      const real* a = W[F[f][0]]; const real* b = W[F[f][1]]; const real* c = W[F[f][2]];
      real lf[3], p[3];
      b2n_closest_triangle(a, b, c, lf);
      for (i = 0; i < 3; i++) p[i] = lf[0] * a[i] + lf[1] * b[i] + lf[2] * c[i];
      const real dd = B2N_DOT(p, p);
      if (dd < best) {
        best = dd;
        l[0] = l[1] = l[2] = l[3] = 0;
        l[F[f][0]] = lf[0]; l[F[f][1]] = lf[1]; l[F[f][2]] = lf[2];
      }
    }
    return 0;
  }
}

/* GJK on the cores.  Returns 1 with the closest points pa / pb and their distance when the cores are apart; 0 when they
 * touch or overlap, leaving the last simplex (Minkowski points W, support points on A in SA, n_simplex) for EPA.         */
B2N_FN int b2n_gjk(const b2n_shape* A, const b2n_shape* B, B2N_REAL* pa, B2N_REAL* pb, B2N_REAL* dist, B2N_REAL W[4][3],
                   B2N_REAL SA[4][3], int* n_simplex) {
  typedef B2N_REAL real;
  real v[3] = {A->c[0] - B->c[0], A->c[1] - B->c[1], A->c[2] - B->c[2]}, l[4] = {1, 0, 0, 0};
  int n = 0, it, i, k;
  if (B2N_DOT(v, v) < (real)1e-12) { v[0] = 1; v[1] = 0; v[2] = 0; }
  for (it = 0; it < B2N_GJK_ITERS; it++) {
    real d[3] = {-v[0], -v[1], -v[2]}, sa[3], sb[3], w[3];
    b2n_support(A, d, sa);
    b2n_support(B, v, sb);
    for (k = 0; k < 3; k++) w[k] = sa[k] - sb[k];
    const real vv = B2N_DOT(v, v), vw = B2N_DOT(v, w);
    if (n > 0 && vv - vw <= (real)1e-5 * vv) break;       /* no point of A - B is closer along v */
    {
      int dup = 0;
      for (i = 0; i < n; i++) {
        real e[3] = {w[0] - W[i][0], w[1] - W[i][1], w[2] - W[i][2]};
        if (B2N_DOT(e, e) <= (real)1e-14) dup = 1;
      }
      if (dup) break;
    }
    if (n == 4) break;   /* cannot happen after a reduction; guards the arrays */
    real Wb[4][3], SAb[4][3], lb[4];   /* the simplex so far: restored if the new vertex brings no progress (rounding) */
    const int nb = n;
    for (i = 0; i < 4; i++) {
      lb[i] = l[i];
      for (k = 0; k < 3; k++) { Wb[i][k] = W[i][k]; SAb[i][k] = SA[i][k]; }
    }
    for (k = 0; k < 3; k++) { W[n][k] = w[k]; SA[n][k] = sa[k]; }
    n++;
    if (b2n_closest_simplex(W, n, l)) { *n_simplex = n; return 0; }
    {
      /* drop the vertices with zero weight, recompute v */
      int m = 0;
      real nv[3] = {0, 0, 0};
      for (i = 0; i < n; i++)
        if (l[i] > 0) {
          for (k = 0; k < 3; k++) { W[m][k] = W[i][k]; SA[m][k] = SA[i][k]; nv[k] += l[i] * W[i][k]; }
          l[m] = l[i];
          m++;
        }
      n = m;
      const real nvv = B2N_DOT(nv, nv);
      if (nvv <= (real)1e-10) { *n_simplex = n; return 0; }                 /* cores touch */
      if (it > 0 && nvv >= vv) {                                            /* no progress (rounding): keep the previous simplex */
        n = nb;
        for (i = 0; i < 4; i++) {
          l[i] = lb[i];
          for (k = 0; k < 3; k++) { W[i][k] = Wb[i][k]; SA[i][k] = SAb[i][k]; }
        }
        break;
      }
      v[0] = nv[0]; v[1] = nv[1]; v[2] = nv[2];
    }
  }
  {
    real a[3] = {0, 0, 0};
    for (i = 0; i < n; i++)
      for (k = 0; k < 3; k++) a[k] += l[i] * SA[i][k];
    /* v was recomputed from the same weights unless the loop stopped on "no progress"; the witness on B follows from it */
    real vs[3] = {0, 0, 0};
    for (i = 0; i < n; i++)
      for (k = 0; k < 3; k++) vs[k] += l[i] * W[i][k];
    for (k = 0; k < 3; k++) { pa[k] = a[k]; pb[k] = a[k] - vs[k]; }
    *dist = B2N_SQRT(B2N_DOT(vs, vs));
  }
  *n_simplex = n;
  return 1;
}

/* EPA from the simplex GJK ended with.  Returns 1 with the unit normal nf of the nearest boundary face of A - B (pointing
 * away from the origin), the depth and the witness points on the cores; 0 if no polytope could be built (cores touching in
 * a point: depth 0).                                                                                                      */
B2N_FN int b2n_epa(const b2n_shape* A, const b2n_shape* B, B2N_REAL W4[4][3], B2N_REAL SA4[4][3], int n, B2N_REAL* nf,
                   B2N_REAL* depth, B2N_REAL* pa, B2N_REAL* pb) {
  typedef B2N_REAL real;
  real V[B2N_EPA_MAXV][3], VA[B2N_EPA_MAXV][3];
  unsigned char F[B2N_EPA_MAXF][3], E[B2N_EPA_MAXE][2];
  int nv = 0, nfc = 0, i, k, it;
  for (i = 0; i < n; i++)
    for (k = 0; k < 3; k++) { V[i][k] = W4[i][k]; VA[i][k] = SA4[i][k]; }
  nv = n;
  /* complete the simplex to a tetrahedron with support points off its affine hull */
  for (it = 0; it < 8 && nv < 4; it++) {
    real d[3] = {0, 0, 0}, md[3], sa[3], sb[3], w[3];
    if (nv == 1) { d[it % 3] = (it & 1) ? (real)-1 : (real)1; }
    else if (nv == 2) {
      real e[3] = {V[1][0] - V[0][0], V[1][1] - V[0][1], V[1][2] - V[0][2]}, ax[3] = {0, 0, 0};
      int mi = B2N_FABS(e[0]) <= B2N_FABS(e[1]) ? (B2N_FABS(e[0]) <= B2N_FABS(e[2]) ? 0 : 2) : (B2N_FABS(e[1]) <= B2N_FABS(e[2]) ? 1 : 2);
      ax[(mi + it) % 3] = 1;
      B2N_CROSS(d, e, ax);
      if (it & 1) { d[0] = -d[0]; d[1] = -d[1]; d[2] = -d[2]; }
    } else {
      real e1[3] = {V[1][0] - V[0][0], V[1][1] - V[0][1], V[1][2] - V[0][2]}, e2[3] = {V[2][0] - V[0][0], V[2][1] - V[0][1], V[2][2] - V[0][2]};
      B2N_CROSS(d, e1, e2);
      if (it & 1) { d[0] = -d[0]; d[1] = -d[1]; d[2] = -d[2]; }
    }
    if (B2N_DOT(d, d) <= (real)1e-30) continue;
    md[0] = -d[0]; md[1] = -d[1]; md[2] = -d[2];
    b2n_support(A, d, sa);
    b2n_support(B, md, sb);
    for (k = 0; k < 3; k++) w[k] = sa[k] - sb[k];
    {
      /* accept only if it enlarges the hull: distinct from the vertices and off the line / plane */
      int ok = 1;
      for (i = 0; i < nv; i++) {
        real e[3] = {w[0] - V[i][0], w[1] - V[i][1], w[2] - V[i][2]};
        if (B2N_DOT(e, e) <= (real)1e-14) ok = 0;
      }
      if (ok && nv >= 2) {
        real e[3] = {w[0] - V[0][0], w[1] - V[0][1], w[2] - V[0][2]};
        const real off = B2N_DOT(e, d);
        if (off * off <= (real)1e-12 * B2N_DOT(d, d)) ok = 0;
      }
      if (!ok) continue;
    }
    for (k = 0; k < 3; k++) { V[nv][k] = w[k]; VA[nv][k] = sa[k]; }
    nv++;
  }
  if (nv < 4) return 0;
  {
    /* orient the four faces outwards (away from the opposite vertex) */
    const int T[4][4] = {{0, 1, 2, 3}, {0, 3, 1, 2}, {0, 2, 3, 1}, {1, 3, 2, 0}};
    for (i = 0; i < 4; i++) {
      const real* a = V[T[i][0]]; const real* b = V[T[i][1]]; const real* c = V[T[i][2]]; const real* o = V[T[i][3]];
      real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, nn[3];
      real ao[3] = {o[0] - a[0], o[1] - a[1], o[2] - a[2]};
      B2N_CROSS(nn, ab, ac);
      F[nfc][0] = (unsigned char)T[i][0];
      if (B2N_DOT(nn, ao) > 0) { F[nfc][1] = (unsigned char)T[i][2]; F[nfc][2] = (unsigned char)T[i][1]; }
      else { F[nfc][1] = (unsigned char)T[i][1]; F[nfc][2] = (unsigned char)T[i][2]; }
      nfc++;
    }
  }
  {
    int bf = 0;
    real bn[3] = {0, 0, 1}, bd = 0;
    for (it = 0;; it++) {
      /* nearest face */
      bd = (real)1e30; bf = -1;
      for (i = 0; i < nfc; i++) {
        const real* a = V[F[i][0]]; const real* b = V[F[i][1]]; const real* c = V[F[i][2]];
        real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, nn[3];
        B2N_CROSS(nn, ab, ac);
        const real l2 = B2N_DOT(nn, nn);
        if (l2 <= (real)1e-30) continue;
        const real il = 1 / B2N_SQRT(l2);
        real dd = B2N_DOT(nn, a) * il;
        if (dd < 0) dd = 0;   /* origin marginally outside (rounding) */
        if (dd < bd) { bd = dd; bf = i; bn[0] = nn[0] * il; bn[1] = nn[1] * il; bn[2] = nn[2] * il; }
      }
      if (bf < 0) return 0;
      if (it >= B2N_EPA_ITERS || nv >= B2N_EPA_MAXV || nfc + 2 > B2N_EPA_MAXF) break;
      real mb[3] = {-bn[0], -bn[1], -bn[2]}, sa[3], sb[3], w[3];
      b2n_support(A, bn, sa);
      b2n_support(B, mb, sb);
      for (k = 0; k < 3; k++) w[k] = sa[k] - sb[k];
      if (B2N_DOT(w, bn) - bd <= (real)1e-6 + (real)1e-4 * bd) break;   /* the face is on the boundary */
      /* remove the faces seen from w, collect the horizon */
      int ne = 0, j, e, found, overflow = 0;
      for (i = 0; i < nfc;) {
        const real* a = V[F[i][0]]; const real* b = V[F[i][1]]; const real* c = V[F[i][2]];
        real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, nn[3];
        real aw[3] = {w[0] - a[0], w[1] - a[1], w[2] - a[2]};
        B2N_CROSS(nn, ab, ac);
        if (B2N_DOT(nn, aw) > 0) {
          for (j = 0; j < 3; j++) {
            const unsigned char ea = F[i][j], eb = F[i][(j + 1) % 3];
            found = 0;
            for (e = 0; e < ne; e++)
              if (E[e][0] == eb && E[e][1] == ea) { E[e][0] = E[ne - 1][0]; E[e][1] = E[ne - 1][1]; ne--; found = 1; break; }
            if (!found) {
              if (ne < B2N_EPA_MAXE) { E[ne][0] = ea; E[ne][1] = eb; ne++; }
              else overflow = 1;
            }
          }
          F[i][0] = F[nfc - 1][0]; F[i][1] = F[nfc - 1][1]; F[i][2] = F[nfc - 1][2];
          nfc--;
        } else i++;
      }
      if (overflow || ne == 0 || nfc + ne > B2N_EPA_MAXF) break;   /* keep the answer of the last complete polytope */
      for (k = 0; k < 3; k++) { V[nv][k] = w[k]; VA[nv][k] = sa[k]; }
      for (e = 0; e < ne; e++) { F[nfc][0] = E[e][0]; F[nfc][1] = E[e][1]; F[nfc][2] = (unsigned char)nv; nfc++; }
      nv++;
    }
    if (bf >= nfc) {   /* the loop left after removing faces: take the nearest face of what remains */
      bd = (real)1e30; bf = -1;
      for (i = 0; i < nfc; i++) {
        const real* a = V[F[i][0]]; const real* b = V[F[i][1]]; const real* c = V[F[i][2]];
        real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, nn[3];
        B2N_CROSS(nn, ab, ac);
        const real l2 = B2N_DOT(nn, nn);
        if (l2 <= (real)1e-30) continue;
        const real il = 1 / B2N_SQRT(l2);
        real dd = B2N_DOT(nn, a) * il;
        if (dd < 0) dd = 0;
        if (dd < bd) { bd = dd; bf = i; bn[0] = nn[0] * il; bn[1] = nn[1] * il; bn[2] = nn[2] * il; }
      }
      if (bf < 0) return 0;
    }
    {
      /* witness: barycentric coordinates of the foot of the origin on the face */
      const int ia = F[bf][0], ib = F[bf][1], ic = F[bf][2];
      real a[3], b[3], c[3], l[3];
      for (k = 0; k < 3; k++) { a[k] = V[ia][k] - bd * bn[k]; b[k] = V[ib][k] - bd * bn[k]; c[k] = V[ic][k] - bd * bn[k]; }
      b2n_closest_triangle(a, b, c, l);
      for (k = 0; k < 3; k++) {
        pa[k] = l[0] * VA[ia][k] + l[1] * VA[ib][k] + l[2] * VA[ic][k];
        const real pk = l[0] * V[ia][k] + l[1] * V[ib][k] + l[2] * V[ic][k];
        pb[k] = pa[k] - pk;
      }
      nf[0] = bn[0]; nf[1] = bn[1]; nf[2] = bn[2];
      *depth = bd;
    }
  }
  return 1;
}

/* One contact of two rounded convex shapes: 1 if their surfaces are closer than `margin` (n from B towards A, points on the
 * two surfaces, dist < 0 = penetration).  id: 0 = cores apart (GJK), 1 = cores overlapping (EPA).                            */
B2N_FN int b2n_convex_contact(const b2n_shape* A, const b2n_shape* B, B2N_REAL margin, B2N_REAL* n_out, b2n_contact* out) {
  typedef B2N_REAL real;
  real W[4][3], SA[4][3], pa[3], pb[3], d = 0, n[3] = {0, 0, 1};
  int ns = 0, k, id = 0;
  if (b2n_gjk(A, B, pa, pb, &d, W, SA, &ns)) {
    if (!(d - A->r - B->r < margin)) return 0;
    for (k = 0; k < 3; k++) n[k] = (pa[k] - pb[k]) / d;
  } else {
    real nfc[3], depth = 0;
    id = 1;
    if (b2n_epa(A, B, W, SA, ns, nfc, &depth, pa, pb)) {
      for (k = 0; k < 3; k++) n[k] = -nfc[k];
      d = -depth;
    } else {
      /* touching in a point: no direction from the cores; separate along the line of centres */
      real cc[3] = {A->c[0] - B->c[0], A->c[1] - B->c[1], A->c[2] - B->c[2]};
      const real l = B2N_SQRT(B2N_DOT(cc, cc));
      if (l > (real)1e-9) { n[0] = cc[0] / l; n[1] = cc[1] / l; n[2] = cc[2] / l; }
      real a[3] = {0, 0, 0};
      for (k = 0; k < 3; k++) a[k] = ns > 0 ? SA[0][k] : A->c[k];
      for (k = 0; k < 3; k++) { pa[k] = a[k]; pb[k] = a[k]; }
      d = 0;
    }
  }
  for (k = 0; k < 3; k++) {
    out->pa[k] = pa[k] - n[k] * A->r;
    out->pb[k] = pb[k] + n[k] * B->r;
    n_out[k] = n[k];
  }
  out->dist = d - A->r - B->r;
  out->id = id;
  return 1;
}

#endif /* B2ENV_NARROWPHASE_H */
