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
    const real tol = (real)1e-5;
    int b0 = 0;
    for (i = 1; i < m; i++)
      if (ph[i] < ph[b0] - tol || (B2N_FABS(ph[i] - ph[b0]) <= tol && lab[i] < lab[b0])) b0 = i;
    int b1 = -1; real v1 = -1;
    for (i = 0; i < m; i++) {
      if (i == b0) continue;
      real d2 = (px[i] - px[b0]) * (px[i] - px[b0]) + (py[i] - py[b0]) * (py[i] - py[b0]);
      if (b1 < 0 || d2 > v1 + tol * tol || (B2N_FABS(d2 - v1) <= tol * tol && lab[i] < lab[b1])) { b1 = i; v1 = d2; }
    }
    const real ex = px[b1] - px[b0], ey = py[b1] - py[b0];
    int b2 = -1, b3 = -1; real v2 = tol * B2N_SQRT(v1 > 0 ? v1 : 0), v3 = v2;
    for (i = 0; i < m; i++) {
      if (i == b0 || i == b1) continue;
      real cr = ex * (py[i] - py[b0]) - ey * (px[i] - px[b0]);
      if (cr > 0) { if (b2 < 0 ? cr > v2 : (cr > v2 + tol * tol || (B2N_FABS(cr - v2) <= tol * tol && lab[i] < lab[b2]))) { b2 = i; v2 = cr; } }
      else { if (b3 < 0 ? -cr > v3 : (-cr > v3 + tol * tol || (B2N_FABS(-cr - v3) <= tol * tol && lab[i] < lab[b3]))) { b3 = i; v3 = -cr; } }
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

#endif /* B2ENV_NARROWPHASE_H */
