// b2env_tree.cuh — the fused step kernel for kinematic TREES of up to 32 bodies (the iCub: 32 revolute joints
// after the six welded F/T-sensor links are folded into their parents), included by b2env.cu.
//
// One env.step() of the reference's iCub task envs (icub_envs/icub_push_gym_env.py:271-282,
// icub_reach_gym_env.py:248-259): hand-pose increment + clamps (:227-253) -> damped-least-squares IK
// (icub_env.py:303-317) -> 32 position motors (:319-326) -> p.stepSimulation (:263) -> observation
// (:166-203, icub_env.py:202-249) -> termination (:327-342) -> reward (:344-373), one launch.
//
// Mapping: ONE WARP = ONE ENVIRONMENT, lane = body = dof (a body is the child of its movable joint, so
// body index == dof index; fixed joints are merged on the host).  The statement is the one of the 16-lane
// Panda kernel: world-frame composite-rigid-body inertia matrix, Gauss-Jordan inverse held one
// row per lane, Delassus-form projected Gauss-Seidel with lane = constraint row (32 motor rows = one row per
// lane, then up to 32 "generic" rows: joint limits and contact normals / frictions).

#ifndef TREE_WPB
#define TREE_WPB 4      // warps (= environments) per block
#endif
#ifndef TREE_MINB
#define TREE_MINB 3     // blocks per SM: 12 warps (168 registers, 17.4 KB of shared memory per env)
#endif
#define TREE_MAXC 8     // contacts kept per env (4 cube-table + 4 proxy contacts): 3 + 3*8 = 27 generic rows <= 32
#define TREE_WS 41      // row stride of the W table (32 arm dofs + 6 cube components; odd: rows AND columns are conflict-free)
#define TREE_ROUNDS 24  // child -> parent accumulation rounds the host schedule may use
#define SHW(v, src) __shfl_sync(FULL, (v), (src))
#ifndef TREE_PHASE_SYNC
#define TREE_PHASE_SYNC 1   // block barriers between the stages (and per IK iteration): the warps of a block walk the
#endif                      // code together and share instruction-cache lines (the kernel is instruction-fetch bound)
#ifndef TREE_BAR_POST_SOLVE
#define TREE_BAR_POST_SOLVE 1
#endif
#ifndef TREE_BAR_IK_ITER
#define TREE_BAR_IK_ITER 1
#endif
#if TREE_PHASE_SYNC
#define TREE_BARRIER() __syncthreads()
#else
#define TREE_BARRIER() __syncwarp()
#endif

struct TreeSmem {
  float Mi[32 * 32];         // joint-space inertia, then its inverse, SKEWED: element (r, c) at r*32 + ((r-c) & 31) —
                             // rows and columns are both conflict-free, rows are 16-byte aligned (see MI())
  float A[32 * 32];          // generic x generic Delassus block, A[c*32 + r]
  float W[32 * TREE_WS];     // W[g][k] = (M^-1 J_g^T)_k; k < 32 arm dofs, 32..37 cube (lin, ang)
  float T[32][12];           // body world transforms: R (9) + p (3)
  float S[32][6];            // world spatial axes about O = base position: (w ; v_O)
  float vstar[44];           // unconstrained velocities (32 arm + 6 cube)
  float mlam[32];            // motor-row impulses
  float glam[32];            // generic-row impulses
  float scr[200];            // IK: Jacobian rows (stride 33); later the observation staging
  Contact con[TREE_MAXC];
  float con_cfm[TREE_MAXC];
  int lim_d[4];
  float lim_dist[4];
  int ckey[B2E_CACHE_SLOTS];
  float clam[B2E_CACHE_SLOTS][3];
};
static_assert(sizeof(TreeSmem) % 16 == 0 && offsetof(TreeSmem, Mi) % 16 == 0 && offsetof(TreeSmem, mlam) % 16 == 0,
              "TreeSmem: Mi and mlam are read with 16-byte loads");
#define MI(r, c) ((r) * 32 + (((r) - (c)) & 31))

__device__ __forceinline__ float wmaxf(float v) {   // max of non-negative floats over the warp
  return __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(v)));
}
__device__ __forceinline__ float wsumf(float v) {   // butterfly sum: every lane gets the same bits
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// forward kinematics, lane = body: local joint transform, composition along the tree by pointer jumping
__device__ __forceinline__ void tree_fk(const DevModel* __restrict__ M, const DevModelU& U, int lane, float qi, float* R,
                                        float* p) {
  const int nl = U.n_links;
  const bool act = lane < nl;
  const int li = act ? lane : 0;
  float jr[9], ax[3];
#pragma unroll
  for (int k = 0; k < 9; k++) jr[k] = __ldg(&M->jrot[li][k]);
#pragma unroll
  for (int k = 0; k < 3; k++) { ax[k] = __ldg(&M->axis[li][k]); p[k] = __ldg(&M->jpos[li][k]); }
  const int jt = __ldg(&M->jtype[li]);
  if (jt == B2E_JOINT_REVOLUTE) {
    float s, c;
    sincosf(qi, &s, &c);
    const float t = 1 - c, x = ax[0], y = ax[1], z = ax[2];
    const float Rq[9] = {t * x * x + c,     t * x * y - s * z, t * x * z + s * y,
                         t * x * y + s * z, t * y * y + c,     t * y * z - s * x,
                         t * x * z - s * y, t * y * z + s * x, t * z * z + c};
    m3mul(jr, Rq, R);
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = jr[k];
    if (jt == B2E_JOINT_PRISMATIC) {
      const float t[3] = {ax[0] * qi, ax[1] * qi, ax[2] * qi};
      float o[3];
      m3vec(jr, t, o);
      p[0] += o[0]; p[1] += o[1]; p[2] += o[2];
    }
  }
  int anc = act ? __ldg(&M->parent[li]) : -1;
  const int rounds = U.fk_rounds;
  for (int rd = 0; rd < rounds; rd++) {
    const int src = anc < 0 ? 0 : anc;
    float Ra[9], pa[3];
#pragma unroll
    for (int k = 0; k < 9; k++) Ra[k] = SHW(R[k], src);
#pragma unroll
    for (int k = 0; k < 3; k++) pa[k] = SHW(p[k], src);
    const int anca = SHW(anc, src);
    if (anc >= 0) {
      float Rn[9], o[3];
      m3mul(Ra, R, Rn);
      m3vec(Ra, p, o);
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = Rn[k];
      p[0] = pa[0] + o[0]; p[1] = pa[1] + o[1]; p[2] = pa[2] + o[2];
      anc = anca;
    }
  }
  {  // base pose
    float Rb[9], Rn[9], o[3];
#pragma unroll
    for (int k = 0; k < 9; k++) Rb[k] = U.base_rot[k];
    m3mul(Rb, R, Rn);
    m3vec(Rb, p, o);
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = Rn[k];
    p[0] = U.base_pos[0] + o[0]; p[1] = U.base_pos[1] + o[1]; p[2] = U.base_pos[2] + o[2];
  }
}

// inclusive sum over the path base..body of a 6-vector held per body lane (pointer jumping)
__device__ __forceinline__ void tree_path_sum6(const DevModel* __restrict__ M, const DevModelU& U, int lane, float* x) {
  int anc = lane < U.n_links ? __ldg(&M->parent[lane]) : -1;
  const int rounds = U.fk_rounds;
  for (int rd = 0; rd < rounds; rd++) {
    const int src = anc < 0 ? 0 : anc;
    float xa[6];
#pragma unroll
    for (int k = 0; k < 6; k++) xa[k] = SHW(x[k], src);
    const int anca = SHW(anc, src);
    if (anc >= 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) x[k] += xa[k];
      anc = anca;
    }
  }
}

// Damped-least-squares IK (p.calculateInverseKinematics with jointDamping, icub_env.py:307-312): the statement
// of oracle/b2oracle.c ik_dls.  Every joint on the base -> hand path is a controlled joint with damping 0.1 and
// the blocked joints (damping 100) have zero Jacobian columns, so (J^T J + D)^-1 J^T e = J^T (J J^T + 0.1 I)^-1 e.
// lane = dof for the Jacobian column and dq; lanes 0..5 hold the rows of the 6x6 system.
__device__ __noinline__ float tree_ik(float* scr, const DevModel* __restrict__ M, const DevModelU& U, int max_iters,
                                      float residual, float damping, int lane, float my_q, float tpx, float tpy, float tpz,
                                      float tqx, float tqy, float tqz, float tqw) {
  const int nl = U.n_links, ee = U.ee_link;
  const int li = lane < nl ? lane : 0;
  const bool on_path = lane < nl && ((U.ee_dofmask >> lane) & 1u);
  const int jt = __ldg(&M->jtype[li]);
  const float ax[3] = {__ldg(&M->axis[li][0]), __ldg(&M->axis[li][1]), __ldg(&M->axis[li][2])};
  float qv = my_q;
#if TREE_PHASE_SYNC && TREE_BAR_IK_ITER
  bool done = false;
  for (int it = 0;; it++) {
    // the loop is shared by the block: every warp takes part in the barrier of every pass until all are done
    if (!__syncthreads_or(!done && it < max_iters)) break;
    if (done || it >= max_iters) continue;
#else
  for (int it = 0; it < max_iters; it++) {
#endif
    LANE_SITE(6);
    float R[9], p[3];
    tree_fk(M, U, lane, lane < nl ? qv : 0.f, R, p);
    float pe[3], Re[9];
#pragma unroll
    for (int k = 0; k < 3; k++) pe[k] = SHW(p[k], ee);
#pragma unroll
    for (int k = 0; k < 9; k++) Re[k] = SHW(R[k], ee);
    const float dp[3] = {tpx - pe[0], tpy - pe[1], tpz - pe[2]};
#if TREE_PHASE_SYNC && TREE_BAR_IK_ITER
    if (sqrtf(dot3(dp, dp)) <= residual) { done = true; continue; }   // warp-uniform
#else
    if (sqrtf(dot3(dp, dp)) <= residual) break;   // warp-uniform
#endif
    float cq[4], eq[4], er[3];
    mat_to_quat(Re, cq);
    {
      const float tq[4] = {tqx, tqy, tqz, tqw}, cc[4] = {-cq[0], -cq[1], -cq[2], cq[3]};
      quat_mul(tq, cc, eq);
    }
    if (eq[3] < 0) { eq[0] = -eq[0]; eq[1] = -eq[1]; eq[2] = -eq[2]; eq[3] = -eq[3]; }
    const float vn = sqrtf(eq[0] * eq[0] + eq[1] * eq[1] + eq[2] * eq[2]);
    if (vn > 1e-9f) {
      const float ang = 2.f * atan2f(vn, eq[3]);
      er[0] = eq[0] / vn * ang; er[1] = eq[1] / vn * ang; er[2] = eq[2] / vn * ang;
    } else { er[0] = er[1] = er[2] = 0.f; }
    float c[6] = {0, 0, 0, 0, 0, 0};   // Jacobian column of this joint
    if (on_path) {
      float aw[3];
      m3vec(R, ax, aw);
      if (jt == B2E_JOINT_REVOLUTE) {
        const float rel[3] = {pe[0] - p[0], pe[1] - p[1], pe[2] - p[2]};
        cross3(aw, rel, c);
        c[3] = aw[0]; c[4] = aw[1]; c[5] = aw[2];
      } else {
        c[0] = aw[0]; c[1] = aw[1]; c[2] = aw[2];
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 6; k++) scr[k * 33 + lane] = c[k];
    __syncwarp();
    const int r = lane < 6 ? lane : 0;   // lane r < 6: row r of J J^T + lambda I, augmented with e_r
    float Ur[7];
#pragma unroll
    for (int cc2 = 0; cc2 < 6; cc2++) Ur[cc2] = (cc2 == r) ? damping : 0.f;
    // only the joints between the base and the hand have a Jacobian column (the others are exact zeros: skipping
    // them leaves every partial sum unchanged); row stride 33: the six row lanes read six different banks
    for (unsigned pm = U.ee_dofmask; pm; pm &= pm - 1) {
      const int d = __ffs(pm) - 1;
      const float cr = scr[r * 33 + d];
#pragma unroll
      for (int cc2 = 0; cc2 < 6; cc2++) Ur[cc2] = fmaf(cr, scr[cc2 * 33 + d], Ur[cc2]);
    }
    Ur[6] = r < 3 ? (r == 0 ? dp[0] : (r == 1 ? dp[1] : dp[2])) : (r == 3 ? er[0] : (r == 4 ? er[1] : er[2]));
#pragma unroll
    for (int k = 0; k < 6; k++) {
      float rk[7];
#pragma unroll
      for (int j = 0; j < 7; j++) rk[j] = SHW(Ur[j], k);
      const float pinv = 1.0f / rk[k];
      if (lane == k) {
#pragma unroll
        for (int j = 0; j < 7; j++) Ur[j] = rk[j] * pinv;
      } else {
        const float f = Ur[k] * pinv;
#pragma unroll
        for (int j = 0; j < 7; j++) Ur[j] = fmaf(-f, rk[j], Ur[j]);
      }
    }
    float dq = 0.f;   // dq_d = sum_r J[r][d] x_r
#pragma unroll
    for (int rr = 0; rr < 6; rr++) dq = fmaf(c[rr], SHW(Ur[6], rr), dq);
    const float mx = wmaxf(fabsf(dq));
    const float scale = mx > 0.78539816339f ? 0.78539816339f / mx : 1.f;
    qv = fmaf(dq, scale, qv);
  }
  __syncwarp();
  return qv;
}

struct TreeRow {   // the generic row of this lane
  float lam, u, base, gg, invd, diag, lo, hi, mu, prev;
  int type, isl, nidx;
};

struct AffineLim {   // the (up to 3) joint-limit rows of the arm island, lane-uniform except lrow
  int ld[3];
  float ls[3], lrow[3], lidg[3], ldiag[3], lrhs[3], x10, x20, x21;
};

// The sweeps of tree_arm_affine, specialised on the number of limit rows: p' = G p + c with p read back as
// lane-uniform 16-byte loads from shared memory (1 STS + 8 LDS.128 instead of 31 shuffles), then the limit rows;
// a motor force bound that would activate is folded into the residual reduction as +inf (one REDUX per sweep).
template <int NL>
__device__ __forceinline__ int tree_affine_sweeps(float* pb, int lane, bool row, const float (&G)[32], float c, float diag,
                                                  float lo, float hi, const AffineLim& L, int max_iters, float tol,
                                                  float& lam_m, float* lam_l) {
  float p = 0.f, lm = 0.f, ll[3] = {0.f, 0.f, 0.f};
  int it = 0;
  bool clamp = false;
  const float INF = __int_as_float(0x7f800000);
  for (; it < max_iters; it++) {
    LANE_SITE(5);
    __syncwarp();
    pb[lane] = p;
    __syncwarp();
    float pk[32];
#pragma unroll
    for (int q4 = 0; q4 < 8; q4++) {
      const float4 v = *reinterpret_cast<const float4*>(pb + 4 * q4);
      pk[4 * q4] = v.x; pk[4 * q4 + 1] = v.y; pk[4 * q4 + 2] = v.z; pk[4 * q4 + 3] = v.w;
    }
    float a0 = c, a1 = 0.f;
#pragma unroll
    for (int k = 1; k < 32; k += 2) a0 = fmaf(G[k], pk[k], a0);
#pragma unroll
    for (int k = 2; k < 32; k += 2) a1 = fmaf(G[k], pk[k], a1);
    float p1 = a0 + a1;                      // G[0] = 0: nothing of the previous iterate enters row 0 from before it
    const float dm = row ? p1 - p : 0.f;     // impulse change of this lane's motor row
    if (!row) p1 = 0.f;
    const float nlm = lm + dm;
    const bool cl = row && !(nlm >= lo && nlm <= hi);   // a motor force bound would activate: not affine any more
    float rv = dm * diag;
    rv = rv * rv;
    if (NL > 0) {
      // limit rows, one after the other, with their exact projection lambda >= 0 (a speculative row of a joint that
      // is merely close to its limit stays at 0).  (M^-1 p)_d of the rows: one interleaved butterfly, then the
      // scalar corrections for the rows handled before.
      float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int l = 0; l < NL; l++) d[l] = L.lrow[l] * p1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int l = 0; l < NL; l++) d[l] += __shfl_xor_sync(FULL, d[l], o);
      }
      float dp[3] = {0.f, 0.f, 0.f};   // generalised impulse changes s_l * dlambda_l
#pragma unroll
      for (int l = 0; l < NL; l++) {
        if (l == 1) d[1] = fmaf(L.x10, dp[0], d[1]);
        if (l == 2) d[2] = fmaf(L.x20, dp[0], fmaf(L.x21, dp[1], d[2]));
        const float nl = fmaxf(fmaf(L.lrhs[l] - L.ls[l] * d[l], L.lidg[l], ll[l]), 0.f);
        const float dl = nl - ll[l];
        ll[l] = nl;
        dp[l] = L.ls[l] * dl;
        const float r2 = dl * L.ldiag[l];
        rv = fmaxf(rv, r2 * r2);
      }
#pragma unroll
      for (int l = 0; l < NL; l++) if (lane == L.ld[l]) p1 += dp[l];
    }
    rv = wmaxf(cl ? INF : rv);
    if (rv == INF) { clamp = true; break; }
    p = p1;
    lm = nlm;
    if (rv <= tol) { it++; break; }
  }
  if (clamp) {
#if defined(B2E_EMU) && defined(EMU_TRACE_AFFINE)
    if (lane == 0) fprintf(stderr, "affine fallback at sweep %d (nlim %d)\n", it, NL);
#endif
    return -1;
  }
#if defined(B2E_EMU) && defined(EMU_TRACE_AFFINE)
  if (lane == 0) fprintf(stderr, "affine ok: %d sweeps (nlim %d)\n", it, NL);
#endif
  lam_m = lm;
#pragma unroll
  for (int l = 0; l < 3; l++) lam_l[l] = ll[l];
  return it;
}

// Arm island made of the position-motor rows and (optionally) joint-limit rows, no arm contact: every row acts
// through one generalised coordinate (J = e_d, or -+e_d for a limit of dof d), so the island's state is the
// generalised impulse p (lane = dof) and one Gauss-Seidel sweep over the 32 motor rows IS the affine map
// p' = G p + c, G = -(D+L)^-1 U, c = (D+L)^-1 b with L + D + U = M^-1 (the motor block of the Delassus matrix);
// a limit row then resets the velocity of its dof: p_d += -+(rhs - -+(M^-1 p)_d) / M^-1_dd.  Same iterates as the
// serial sweep in exact arithmetic, same per-sweep residual test — the 32 serial, shuffle-dependent row updates
// of a sweep become one 32-wide mat-vec.  This matters for the iCub: the IK ignores joint limits
// (icub_env.py:307-312), a target beyond a limit makes the motor row and the limit row of that joint contradict
// each other, and the solver then runs all 150 sweeps (the wrist pitch sits on its limit in the home hand pose).
// Limit rows keep their exact projection (lambda >= 0); if a MOTOR force bound would activate the caller falls back to
// the serial sweep.
// Returns the sweep count, or -1 on fallback; motor impulses in lam_m, limit impulses in lam_l[0..nlim).
__device__ __noinline__ int tree_arm_affine(TreeSmem& sm, int lane, int nd, int nlim, float b, float invd, float diag,
                                            float lo, float hi, float rl0, float rl1, float rl2, int max_iters, float tol,
                                            float& lam_m, float* lam_l) {
  const bool row = lane < nd;
  const float* Mi = sm.Mi;
  float G[32];
  float c = 0.f;
  {
    // G = -(D+L)^-1 U and c = (D+L)^-1 b, lane = row r, in ONE ROLLED loop (no pivot-dependent register index, no
    // shuffle-dependent chain).  Row r of T = (D+L)^-1 solves the upper-triangular system (D+L)^T y = e_r by back
    // substitution in column form: y_j = rhs_j / A_jj, then rhs_i -= A_ji y_j for i < j, j = 31..0; every finished
    // y_j = T[r][j] is consumed at once: G[r][k] -= y_j A_jk for k > j, c_r += y_j b_j.  Both register files rotate one
    // place per step (z[m] = rhs[j-m]; G[m] = G_r[j+1+m] on entry, natural order after j = 0), so all indices
    // are static; slots that rotate in from outside the triangle only ever feed slots outside the triangle.
    // The skewed table IS what this loop wants, X[j][m] = A[j][(j-m) mod 32] (zero outside the n_dof block):
    // m = 1..j is the lower part of row j (the solve), m = j+1..31 the upper part
    // reversed (the G update), read as lane-uniform 16-byte loads.
    const float* X = Mi;
    const float idg = row ? invd : 1.f;
    const float bb = row ? b : 0.f;
    float z[32];
#pragma unroll
    for (int m = 0; m < 32; m++) { z[m] = (31 - m == lane) ? 1.f : 0.f; G[m] = 0.f; }
#pragma unroll 1
    for (int j = 31; j >= 0; j--) {
      float xr[32];
#pragma unroll
      for (int q4 = 0; q4 < 8; q4++) {
        const float4 v = *reinterpret_cast<const float4*>(X + j * 32 + 4 * q4);
        xr[4 * q4] = v.x; xr[4 * q4 + 1] = v.y; xr[4 * q4 + 2] = v.z; xr[4 * q4 + 3] = v.w;
      }
      const float yj = z[0] * SHW(idg, j);
      c = fmaf(yj, SHW(bb, j), c);
#pragma unroll
      for (int m = 1; m < 32; m++) z[m - 1] = fmaf(-yj, xr[m], z[m]);
#pragma unroll
      for (int m = 31; m >= 1; m--) G[m] = fmaf(-yj, xr[32 - m], G[m - 1]);
      G[0] = 0.f;
    }
  }
  // limit rows: dof, sign of J, row of M^-1, diagonal, and the couplings M^-1[d_l][d_l'] between them
  int ld[3] = {0, 0, 0};
  float ls[3] = {0.f, 0.f, 0.f}, lrow[3] = {0.f, 0.f, 0.f}, lidg[3] = {1.f, 1.f, 1.f}, ldiag[3] = {1.f, 1.f, 1.f};
  float x10 = 0.f, x20 = 0.f, x21 = 0.f;
  const float lrhs[3] = {rl0, rl1, rl2};
#pragma unroll
  for (int l = 0; l < 3; l++) {
    if (l < nlim) {
      const int code = sm.lim_d[l];
      ld[l] = code & 0xff;
      ls[l] = (code >> 8) ? -1.f : 1.f;
      lrow[l] = Mi[MI(ld[l], lane)];
      ldiag[l] = Mi[MI(ld[l], ld[l])];
      lidg[l] = 1.0f / ldiag[l];
    }
  }
  if (nlim > 1) x10 = Mi[MI(ld[1], ld[0])];
  if (nlim > 2) { x20 = Mi[MI(ld[2], ld[0])]; x21 = Mi[MI(ld[2], ld[1])]; }
  AffineLim L;
#pragma unroll
  for (int l = 0; l < 3; l++) { L.ld[l] = ld[l]; L.ls[l] = ls[l]; L.lrow[l] = lrow[l]; L.lidg[l] = lidg[l]; L.ldiag[l] = ldiag[l]; L.lrhs[l] = lrhs[l]; }
  L.x10 = x10; L.x20 = x20; L.x21 = x21;
  float* pb = sm.mlam;   // the iterate p, broadcast through shared memory (free until the impulses are stored)
  switch (nlim) {
    case 0: return tree_affine_sweeps<0>(pb, lane, row, G, c, diag, lo, hi, L, max_iters, tol, lam_m, lam_l);
    case 1: return tree_affine_sweeps<1>(pb, lane, row, G, c, diag, lo, hi, L, max_iters, tol, lam_m, lam_l);
    case 2: return tree_affine_sweeps<2>(pb, lane, row, G, c, diag, lo, hi, L, max_iters, tol, lam_m, lam_l);
    default: return tree_affine_sweeps<3>(pb, lane, row, G, c, diag, lo, hi, L, max_iters, tol, lam_m, lam_l);
  }
}

// Projected Gauss-Seidel on the Delassus form: rows in Bullet's order (motors, limits, contact normals, then the
// frictions bounded by mu * normal impulse); the arm island and the cube island converge independently unless a
// proxy-cube contact couples them.  Same iteration as oracle/b2oracle.c physics_step.  Warp-uniform control flow.
__device__ __forceinline__ int tree_pgs(TreeSmem& sm, int lane, MotorRegs& m, TreeRow& r, int nd, int RG, int fric_start,
                                        bool coupled, bool has_cube_rows, bool arm_done, int max_iters, float tol) {
  const bool valid = lane < RG, fr = lane >= fric_start;
  const bool cube = !coupled && r.isl == 1;
  const unsigned arm_nf = __ballot_sync(FULL, valid && !cube && !fr), arm_f = __ballot_sync(FULL, valid && !cube && fr);
  const unsigned cube_nf = __ballot_sync(FULL, valid && cube && !fr), cube_f = __ballot_sync(FULL, valid && cube && fr);
  const float* Mi = sm.Mi;
  bool done0 = arm_done, done1 = !has_cube_rows || coupled;   // arm_done: the arm island was solved by tree_arm_affine
  int it = 0;
  for (; it < max_iters; it++) {
    if (done0 && done1) break;
    LANE_SITE(4);
    m.prev = m.lam;
    r.prev = r.lam;
    r.base = r.lam * r.gg;
    if (!done0) {
      for (int i = 0; i < nd; i++) {   // motor rows: Delassus column i = [M^-1[:, i] ; W_g[i]]
        const float cm = Mi[MI(i, lane)], cg = sm.W[lane * TREE_WS + i];   // column i of W (rows >= RG are zero)
        float nl = fmaf(m.u, m.invd, m.lam);
        nl = fminf(fmaxf(nl, m.lo), m.hi);
        const float dli = SHW(nl - m.lam, i);
        if (lane == i) m.lam = nl;
        m.u = fmaf(-cm, dli, m.u);
        r.u = fmaf(-cg, dli, r.u);
      }
    }
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
      unsigned mk = phase == 0 ? ((done0 ? 0u : arm_nf) | (done1 ? 0u : cube_nf)) : ((done0 ? 0u : arm_f) | (done1 ? 0u : cube_f));
      if (phase == 1 && mk != 0u) {   // friction bounds from the current normal impulses
        const float v = SHW(r.lam, r.nidx);
        if (r.type == ROW_FRICTION) { const float lim = r.mu * v; r.lo = -lim; r.hi = lim; }
      }
      while (mk) {
        const int g = __ffs(mk) - 1;
        mk &= mk - 1;
        const float ca = sm.A[g * 32 + lane], cw = sm.W[g * TREE_WS + lane];
        float nl = fmaf(r.u, r.invd, r.base);
        nl = fminf(fmaxf(nl, r.lo), r.hi);
        const float dli = SHW(nl - r.lam, g);
        if (lane == g) r.lam = nl;
        m.u = fmaf(-cw, dli, m.u);
        r.u = fmaf(-ca, dli, r.u);
      }
    }
    float ra = 0.f, rc = 0.f;
    if (!done0 && lane < nd) {
      const float rv = (m.lam - m.prev) * m.diag;
      ra = rv * rv;
    }
    if (valid) {
      float rv = (r.lam - r.prev) * r.diag;
      rv = rv * rv;
      if (coupled || r.isl == 0) ra = fmaxf(ra, rv); else rc = fmaxf(rc, rv);
    }
    ra = wmaxf(ra);
    rc = wmaxf(rc);
    if (!done0 && ra <= tol) done0 = true;
    if (!done1 && rc <= tol) done1 = true;
    if (done0 && done1) { it++; break; }
  }
  return it;
}

__device__ __forceinline__ void tree_emit(TreeSmem& sm, int slot, int maxc, int key, int type, int link, const float* pA,
                                          const float* pB, const float* n, float dist, float mu, float erp, float cfm) {
  if (slot >= maxc) return;
  Contact& c = sm.con[slot];
  c.key = key; c.type = type; c.link = link; c.link2 = -1;
#pragma unroll
  for (int j = 0; j < 3; j++) { c.pA[j] = pA[j]; c.pB[j] = pB[j]; c.n[j] = n[j]; }
  c.dist = dist; c.mu = mu; c.erp = erp;
  sm.con_cfm[slot] = cfm;
}

// General collision path of the tree kernel: the static world of round 2 (rim and sides of the top slab, table legs, ground
// plane) for the cube and the sphere proxies, family by family in the order and with the keys of oracle collide()
// (b2oracle.c: cube vs table top | cube vs static boxes through b2n_box_box | cube vs plane | sphere vs cube | sphere vs static
// boxes, at most three each | sphere vs plane).  Taken by a warp (= an environment) only when its cube is not wholly over the
// table top or one of its spheres is near anything but the top face: rare, warp-uniform, so the common path keeps the closed
// forms of the caller.  v: this lane's cube vertex (lane < 8); s_*: this lane's sphere (lane < ns) and its sphere-cube result.
// Returns the contact count (<= maxc) | overflow << 8.
__device__ __noinline__ int tree_collide_general(TreeSmem& sm, const DevModel* __restrict__ M, const b2e_params& P, int lane, int ns,
                                                 int maxc, float cpx, float cpy, float cpz, float cqx, float cqy, float cqz, float cqw,
                                                 bool cube_fast, float vx, float vy, float vz, float scx, float scy, float scz,
                                                 float s_r, int s_link, bool sc_hit, float sc_dist, float scnx, float scny,
                                                 float scnz, float scbx, float scby, float scbz) {
  const unsigned lt = (1u << lane) - 1u;
  const float margin = P.contact_margin, ca = P.cube_half, top = P.table_max[2];
  const float cpos[3] = {cpx, cpy, cpz}, cquat[4] = {cqx, cqy, cqz, cqw};
  const float v[3] = {vx, vy, vz}, s_c[3] = {scx, scy, scz};
  const float up[3] = {0.f, 0.f, 1.f};
  const float rb = ca * 1.7320508075688772f;
  const int nsb = P.n_sboxes > 0 ? P.n_sboxes : 1;
  float Rc[9];
  quat_to_mat(cquat, Rc);
  int base = 0;
  {  // cube wholly over the table top: vertex-face manifold
    const bool hit = cube_fast && lane < 8 && (v[2] - top) < margin;
    const unsigned b = __ballot_sync(FULL, hit);
    if (hit) {
      const float pB[3] = {v[0], v[1], top};
      tree_emit(sm, base + __popc(b & lt), maxc, KEY_CUBE_TABLE + lane, CT_CUBE_STATIC, -1, v, pB, up, v[2] - top, P.cube_mu * P.table_mu, P.erp, 0.f);
    }
    base += __popc(b);
  }
  {  // rim of the top slab, legs: lane = static box, general box-box behind a bounding-sphere cull
    int cnt = 0;
    float nrm[3] = {0.f, 0.f, 1.f};
    b2n_contact pts[B2N_MAX_POINTS];
    const int k = lane;
    if (k < nsb && k >= (cube_fast ? 1 : 0)) {
      const float chh[3] = {ca, ca, ca};
      const float ident[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
      float bc[3], bh[3];
      sbox_get(P, k, bc, bh);
      const float d[3] = {cpos[0] - bc[0], cpos[1] - bc[1], cpos[2] - bc[2]};
      if (!(sqrtf(dot3(d, d)) - (rb + sqrtf(dot3(bh, bh))) >= margin)) cnt = b2n_box_box(cpos, Rc, chh, bc, ident, bh, margin, nrm, pts);
    }
    __syncwarp();
    int off = 0, total = 0;
    for (int j = 0; j < nsb; j++) {
      const int cj = __shfl_sync(FULL, cnt, j);
      if (j < lane) off += cj;
      total += cj;
    }
    for (int q = 0; q < cnt; q++)
      tree_emit(sm, base + off + q, maxc, B2E_KEY_CUBE_SBOX + B2N_ID_STRIDE * k + pts[q].id, CT_CUBE_STATIC, -1, pts[q].pa, pts[q].pb, nrm,
                pts[q].dist, P.cube_mu * (P.n_sboxes > 0 ? P.sbox_mu[k] : P.table_mu), P.erp, 0.f);
    base += total;
  }
  {  // ground plane z = 0
    const bool hit = !cube_fast && lane < 8 && v[2] < margin;
    const unsigned b = __ballot_sync(FULL, hit);
    if (hit) {
      const float pB[3] = {v[0], v[1], 0.f};
      tree_emit(sm, base + __popc(b & lt), maxc, KEY_CUBE_PLANE + lane, CT_CUBE_STATIC, -1, v, pB, up, v[2], P.cube_mu * P.plane_mu, P.erp, 0.f);
    }
    base += __popc(b);
  }
  const bool s_world = lane < ns && !(__ldg(&M->sph_flags[lane < ns ? lane : 0]) & 1);
  const float s_mu = lane < ns ? __ldg(&M->sph_mu[lane]) : 0.f, s_cfm = lane < ns ? __ldg(&M->sph_cfm[lane]) : 0.f;
  const float serp = lane < ns ? __ldg(&M->sph_erp[lane]) : -1.f;
  const float s_erp = serp >= 0.f ? serp : P.erp;
  {  // sphere proxies vs the cube (closed form of the caller)
    const bool hit = s_world && sc_hit;
    const unsigned b = __ballot_sync(FULL, hit);
    if (hit) {
      const float n[3] = {scnx, scny, scnz}, pB[3] = {scbx, scby, scbz};
      const float pA[3] = {s_c[0] - n[0] * s_r, s_c[1] - n[1] * s_r, s_c[2] - n[2] * s_r};
      tree_emit(sm, base + __popc(b & lt), maxc, KEY_SPHERE_CUBE + lane, CT_SPHERE_CUBE, s_link, pA, pB, n, sc_dist, P.cube_mu * s_mu, s_erp, s_cfm);
    }
    base += __popc(b);
  }
  {  // sphere proxies vs the static boxes (lane = sphere, at most three boxes each)
    int cnt = 0, kk[3] = {0, 0, 0};
    float nn[3][3], pb[3][3], dd[3] = {0.f, 0.f, 0.f};
    bool tp[3] = {false, false, false};
    if (s_world) {
      for (int k = 0; k < nsb && cnt < 3; k++) {
        float bc[3], bh[3], n[3], pB[3], dist;
        bool is_top;
        sbox_get(P, k, bc, bh);
        const float d[3] = {s_c[0] - bc[0], s_c[1] - bc[1], s_c[2] - bc[2]};
        if (sqrtf(dot3(d, d)) - (s_r + sqrtf(dot3(bh, bh))) >= margin) continue;
        if (!sphere_aabox(s_c, s_r, bc, bh, margin, n, pB, dist, is_top)) continue;
#pragma unroll
        for (int c = 0; c < 3; c++)
          if (c == cnt) {
            kk[c] = k; dd[c] = dist; tp[c] = is_top;
#pragma unroll
            for (int j = 0; j < 3; j++) { nn[c][j] = n[j]; pb[c][j] = pB[j]; }
          }
        cnt++;
      }
    }
    __syncwarp();
    const unsigned b0 = __ballot_sync(FULL, cnt & 1), b1 = __ballot_sync(FULL, cnt & 2);
    const int off = __popc(b0 & lt) + 2 * __popc(b1 & lt), total = __popc(b0) + 2 * __popc(b1);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (c < cnt) {
        const int k = kk[c];
        const float pA[3] = {s_c[0] - nn[c][0] * s_r, s_c[1] - nn[c][1] * s_r, s_c[2] - nn[c][2] * s_r};
        tree_emit(sm, base + off + c, maxc, (k == 0 && tp[c]) ? KEY_SPHERE_TABLE + lane : B2E_KEY_SPHERE_SBOX + 8 * lane + k, CT_SPHERE_STATIC,
                  s_link, pA, pb[c], nn[c], dd[c], (P.n_sboxes > 0 ? P.sbox_mu[k] : P.table_mu) * s_mu, s_erp, s_cfm);
      }
    }
    base += total;
  }
  {  // sphere proxies vs the ground plane
    const float dist = s_c[2] - s_r;
    const bool hit = s_world && dist < margin;
    const unsigned b = __ballot_sync(FULL, hit);
    if (hit) {
      const float pA[3] = {s_c[0], s_c[1], dist}, pB[3] = {s_c[0], s_c[1], 0.f};
      tree_emit(sm, base + __popc(b & lt), maxc, B2E_KEY_SPHERE_PLANE + lane, CT_SPHERE_STATIC, s_link, pA, pB, up, dist, P.plane_mu * s_mu, s_erp, s_cfm);
    }
    base += __popc(b);
  }
  __syncwarp();
  return (base > maxc ? maxc : base) | (base > maxc ? 256 : 0);
}

template <bool IK>
__global__ void __launch_bounds__(32 * TREE_WPB, TREE_MINB)
tree_step_kernel(const DevModel* __restrict__ M, const __grid_constant__ DevModelU U, const __grid_constant__ b2e_params P,
                 DevState st, const float* __restrict__ action, float* __restrict__ obs_out, float* __restrict__ reward_out,
                 float* __restrict__ done_out, int nsub, int mode, int record_contacts, const int* __restrict__ env_ids,
                 int n_ids, int env_offset, int* sched, int seq) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int n_slots = n_ids;
  const int slot = blockIdx.x * TREE_WPB + warp;
  int env;
  if (sched) {
    // Cost-ordered blocks (full-batch physics launches): the warps of a block advance in phase (stage barriers, block-shared
    // IK loop), so a block lasts as long as its slowest environment.  Every environment files itself, at the end of a step,
    // under a class given by its solver sweep count; the next step takes slot -> environment from the class lists, heaviest
    // class first, so that the four environments of a block cost about the same.  Results do not depend on the slot
    // (tests/test_emu_kernels.py::test_icub_cost_ordered_blocks_are_transparent).  Same list layout as the group kernel's.
    const int qp = (seq - 1) & 3, pp = (seq - 1) & 1, B = st.B;
    if (blockIdx.x == 0 && (int)threadIdx.x < NBK) sched[SCHED_CNT((seq + 1) & 3, threadIdx.x)] = 0;   // the next step's counters
    int rem = slot, cls = 0, tot = 0;
    bool found = false;
    for (int b = NBK_MAIN - 1; b >= 0; b--) {
      const int c = sched[SCHED_CNT(qp, b)];
      tot += c;
      if (!found) {
        if (rem < c) { cls = b; found = true; }
        else rem -= c;
      }
    }
    n_slots = tot;
    if (!found) {   // padding warp of the last block: shadow the last environment of the lightest populated class
      for (int b = 0; b < NBK_MAIN; b++) {
        const int c = sched[SCHED_CNT(qp, b)];
        if (c > 0) { cls = b; rem = c - 1; break; }
      }
    }
    env = sched[SCHED_LIST(pp, cls, B) + rem];
  } else {
    const int sc = slot < n_slots ? slot : n_slots - 1;
    env = env_ids ? env_ids[sc] : env_offset + sc;
  }
  const bool live_env = slot < n_slots;   // padding warps of the last block shadow another slot, stores masked
  TreeSmem& sm = reinterpret_cast<TreeSmem*>(smem_raw)[warp];
  const int nd = U.n_dof, nl = U.n_links;   // nl == nd: body index == dof index
  const float dt = P.dt;
  const int maxc = (P.max_contacts > 0 && P.max_contacts < TREE_MAXC) ? P.max_contacts : TREE_MAXC;

#ifdef PROFILE_WARM   // probe build: active lanes of the warp at the stage boundaries -> B2E_F_CONTACTS[env][32..47] (32 = converged)
  int tp_am[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define TREE_PROBE(k) tp_am[k] = __popc(__activemask())
#else
#define TREE_PROBE(k) do { } while (0)
#endif
  // ---- load state ----
  const bool is_dof = lane < nd;
  const int li = is_dof ? lane : 0;
  float my_q = is_dof ? st.q[env * nd + lane] : 0.f;
  float my_qd = is_dof ? st.qd[env * nd + lane] : 0.f;
  float my_target = is_dof ? st.mtarget[env * nd + lane] : 0.f;
  float cpos[3], cquat[4], cv[3], cw[3], target[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    cpos[k] = st.obj_pose[env * 7 + k];
    cv[k] = st.obj_vel[env * 6 + k];
    cw[k] = st.obj_vel[env * 6 + 3 + k];
    target[k] = st.target[env * 3 + k];
  }
#pragma unroll
  for (int k = 0; k < 4; k++) cquat[k] = st.obj_pose[env * 7 + 3 + k];
  int counter = st.counters[env * 2], terminated = st.counters[env * 2 + 1];
  int flags = st.status[env * 4];
  if (lane < B2E_CACHE_SLOTS) {
    sm.ckey[lane] = st.cache_key[env * B2E_CACHE_SLOTS + lane];
#pragma unroll
    for (int j = 0; j < 3; j++) sm.clam[lane][j] = st.cache_lam[(env * B2E_CACHE_SLOTS + lane) * 3 + j];
  }
  // joint lists of the task (icub_env.py:121-143): which action entry drives this dof, which observation
  // entry shows it, whether it is a controlled joint
  int act_idx = -1, obs_idx = -1;
  bool is_ctrl;
  if (P.n_obs_joints > 0) {
    for (int k = 0; k < P.n_ctrl; k++) if (P.ctrl_dof[k] == lane) act_idx = k;
    for (int k = 0; k < P.n_obs_joints; k++) if (P.obs_dof[k] == lane) obs_idx = k;
    is_ctrl = is_dof && ((P.ctrl_mask >> lane) & 1u);
  } else {
    act_idx = lane < P.n_ctrl ? lane : -1;
    obs_idx = is_dof ? lane : -1;
    is_ctrl = lane < P.n_ctrl;
  }
  const int n_qobs = P.n_obs_joints > 0 ? P.n_obs_joints : nd;
  float my_act = 0.f;   // joint mode: the action entry of this dof; IK mode: entry `lane` of the hand-pose increment
  if (mode == B2E_MODE_ACTION) {
    if (IK) { if (lane < P.n_act) my_act = action[env * P.n_act + lane]; }
    else if (act_idx >= 0) my_act = action[env * P.n_act + act_idx];
  }
  float my_hp = (IK && lane < 6) ? st.hand_pose[env * 6 + lane] : 0.f;
  const float my_lower = is_dof ? __ldg(&M->lower[li]) : 0.f, my_upper = is_dof ? __ldg(&M->upper[li]) : 0.f;
  const float my_home = is_dof ? __ldg(&M->home[li]) : 0.f;
  const bool ctrl_gains = !IK && mode != B2E_MODE_HOLD && mode != B2E_MODE_IK_POSE;
  // Cartesian control with a velocity cap (icub_env.py:331-337): every joint keeps its gain, maxVelocity = max_vel
  const bool ik_cap = IK && (mode == B2E_MODE_ACTION || mode == B2E_MODE_IK_POSE) && P.ik_max_vel > 0.f && is_dof;
  const float my_kp = ik_cap ? P.kp_ik_max_vel : ((ctrl_gains && is_ctrl) ? P.kp_ctrl : P.kp_hold);
  int iters = 0, nc = 0, R = 0;
  bool stop = false;
  __syncwarp();

  float Rm[9], pw[3];
  for (int sub = 0;; sub++) {
    // ---- forward kinematics of the current q (start of this sub-step = end of the previous one) ----
    tree_fk(M, U, lane, is_dof ? my_q : 0.f, Rm, pw);
    // ---- termination inside apply_action (icub_push_gym_env.py:266-269), for the previous sub-step ----
    {
      float e3[3];
      {
        const float cm[3] = {U.ee_com[0], U.ee_com[1], U.ee_com[2]};
        float o[3];
        m3vec(Rm, cm, o);
        e3[0] = SHW(pw[0] + o[0], U.ee_link); e3[1] = SHW(pw[1] + o[1], U.ee_link); e3[2] = SHW(pw[2] + o[2], U.ee_link);
      }
      if (sub > 0 && mode == B2E_MODE_ACTION && !stop) {
        float d;
        if (P.task == B2E_TASK_PUSH) {
          const float dd[3] = {cpos[0] - target[0], cpos[1] - target[1], cpos[2] - target[2]};
          d = sqrtf(dot3(dd, dd));
        } else {
          const float dd[3] = {e3[0] - cpos[0], e3[1] - cpos[1], e3[2] - cpos[2]};
          d = sqrtf(dot3(dd, dd));
        }
        if (P.goal_env) { if (counter > P.max_steps) stop = true; else counter++; }
        else if (d <= P.dist_min) { terminated = 1; stop = true; }
        else if (terminated || counter > P.max_steps) stop = true;
        else counter++;
      }
    }
    if (sub >= nsub) break;
    const bool ghost = stop;   // terminated mid-repeat: keep pace, change nothing

    // ---- action -> motor targets ----
    if (!IK && mode == B2E_MODE_ACTION && act_idx >= 0 && !ghost) {   // icub_push_gym_env.py:256-257, icub_env.py:347-361
      my_act *= P.act_scale;
      my_target = fminf(fmaxf(my_q + my_act, my_lower), my_upper);
    }
    if (IK && (mode == B2E_MODE_ACTION || mode == B2E_MODE_IK_POSE)) {
      // Cartesian control (icub_push_gym_env.py:227-253, icub_env.py:261-326): hand pose += scaled action,
      // clamps, COM -> link frame, IK, blocked joints -> rest pose, position targets for all 32 joints
      if (mode == B2E_MODE_ACTION && !ghost && lane < 6) {
        if (lane < 3) {
          my_act *= P.act_scale_pos;
          my_hp = fminf(fmaxf(my_hp + my_act, P.ws_lim[lane][0]), P.ws_lim[lane][1]);
        } else if (P.ik_orientation) {
          my_act *= P.act_scale_rot;
          my_hp = fminf(fmaxf(my_hp + my_act, P.eu_lim[lane - 3][0]), P.eu_lim[lane - 3][1]);
        }
      }
      float tp[3], eu[3], tq[4];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        tp[k] = SHW(my_hp, k);
        const float e = SHW(my_hp, 3 + k);
        eu[k] = P.ik_orientation ? fminf(fmaxf(e, P.eu_lim[k][0]), P.eu_lim[k][1]) : P.home_hand_pose[3 + k];
      }
      tp[2] = fminf(fmaxf(tp[2], P.ws_lim[2][0]), P.ws_lim[2][1]);
      euler_to_quat(eu, tq);
      {
        float Rt[9], o[3];
        const float off[3] = {P.ik_link_offset[0], P.ik_link_offset[1], P.ik_link_offset[2]};
        quat_to_mat(tq, Rt);
        m3vec(Rt, off, o);
        tp[0] += o[0]; tp[1] += o[1]; tp[2] += o[2];
      }
      const float t = tree_ik(sm.scr, M, U, P.ik_iters, P.ik_residual, P.ik_damping, lane, my_q, tp[0], tp[1], tp[2], tq[0],
                              tq[1], tq[2], tq[3]);
      if (is_dof && !ghost) my_target = (P.n_obs_joints > 0 && !is_ctrl) ? my_home : t;
    }
    TREE_PROBE(1);   // after the action / IK stage
    TREE_BARRIER();
    if (is_dof) {
#pragma unroll
      for (int k = 0; k < 9; k++) sm.T[lane][k] = Rm[k];
#pragma unroll
      for (int k = 0; k < 3; k++) sm.T[lane][9 + k] = pw[k];
    }

    // ---- dynamics in world coordinates about O = base position ----
    TREE_PROBE(2);
    float S[6] = {0, 0, 0, 0, 0, 0};
    float Iw[16];   // f(6) | m | h(3) | I_O(6): summed over the subtree below
    {
      const int jt = __ldg(&M->jtype[li]);
      const float ax[3] = {__ldg(&M->axis[li][0]), __ldg(&M->axis[li][1]), __ldg(&M->axis[li][2])};
      float aw[3];
      m3vec(Rm, ax, aw);
      const float rel[3] = {pw[0] - U.base_pos[0], pw[1] - U.base_pos[1], pw[2] - U.base_pos[2]};
      if (is_dof && jt == B2E_JOINT_REVOLUTE) {
        float t[3];
        cross3(rel, aw, t);
        S[0] = aw[0]; S[1] = aw[1]; S[2] = aw[2]; S[3] = t[0]; S[4] = t[1]; S[5] = t[2];
      } else if (is_dof && jt == B2E_JOINT_PRISMATIC) {
        S[3] = aw[0]; S[4] = aw[1]; S[5] = aw[2];
      }
      const float cm[3] = {__ldg(&M->com[li][0]), __ldg(&M->com[li][1]), __ldg(&M->com[li][2])};
      float c[3];
      m3vec(Rm, cm, c);
      c[0] += rel[0]; c[1] += rel[1]; c[2] += rel[2];
      const float ms = is_dof ? __ldg(&M->mass[li]) : 0.f;
      float Ic[9], RI[9], Icw[9];
#pragma unroll
      for (int k = 0; k < 9; k++) Ic[k] = is_dof ? __ldg(&M->inertia[li][k]) : 0.f;
      m3mul(Rm, Ic, RI);
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Icw[3 * i + j] = RI[3 * i] * Rm[3 * j] + RI[3 * i + 1] * Rm[3 * j + 1] + RI[3 * i + 2] * Rm[3 * j + 2];
      const float cc = dot3(c, c);
      Iw[6] = ms;
      Iw[7] = ms * c[0]; Iw[8] = ms * c[1]; Iw[9] = ms * c[2];
      Iw[10] = Icw[0] + ms * (cc - c[0] * c[0]);
      Iw[11] = Icw[1] - ms * c[0] * c[1];
      Iw[12] = Icw[2] - ms * c[0] * c[2];
      Iw[13] = Icw[4] + ms * (cc - c[1] * c[1]);
      Iw[14] = Icw[5] - ms * c[1] * c[2];
      Iw[15] = Icw[8] + ms * (cc - c[2] * c[2]);
    }
    float vJ[6], v[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { vJ[k] = S[k] * my_qd; v[k] = vJ[k]; }
    tree_path_sum6(M, U, lane, v);
    float ab[6];
    {
      float a[3], b[3], c2[3];
      cross3(v, vJ, a);
      cross3(v, vJ + 3, b);
      cross3(v + 3, vJ, c2);
      ab[0] = a[0]; ab[1] = a[1]; ab[2] = a[2];
      ab[3] = b[0] + c2[0]; ab[4] = b[1] + c2[1]; ab[5] = b[2] + c2[2];
    }
    tree_path_sum6(M, U, lane, ab);
    ab[3] -= P.gravity[0]; ab[4] -= P.gravity[1]; ab[5] -= P.gravity[2];   // a_0 = -g
    {
      const float ms = Iw[6];
      const float* h = &Iw[7];
      const float* I = &Iw[10];
      float Ia_n[3], Ia_f[3], Iv_n[3], Iv_f[3], t[3];
      cross3(h, ab + 3, t);
      Ia_n[0] = I[0] * ab[0] + I[1] * ab[1] + I[2] * ab[2] + t[0];
      Ia_n[1] = I[1] * ab[0] + I[3] * ab[1] + I[4] * ab[2] + t[1];
      Ia_n[2] = I[2] * ab[0] + I[4] * ab[1] + I[5] * ab[2] + t[2];
      cross3(h, ab, t);
      Ia_f[0] = ms * ab[3] - t[0]; Ia_f[1] = ms * ab[4] - t[1]; Ia_f[2] = ms * ab[5] - t[2];
      cross3(h, v + 3, t);
      Iv_n[0] = I[0] * v[0] + I[1] * v[1] + I[2] * v[2] + t[0];
      Iv_n[1] = I[1] * v[0] + I[3] * v[1] + I[4] * v[2] + t[1];
      Iv_n[2] = I[2] * v[0] + I[4] * v[1] + I[5] * v[2] + t[2];
      cross3(h, v, t);
      Iv_f[0] = ms * v[3] - t[0]; Iv_f[1] = ms * v[4] - t[1]; Iv_f[2] = ms * v[5] - t[2];
      float x1[3], x2[3], x3[3];
      cross3(v, Iv_n, x1);
      cross3(v + 3, Iv_f, x2);
      cross3(v, Iv_f, x3);
      Iw[0] = Ia_n[0] + x1[0] + x2[0]; Iw[1] = Ia_n[1] + x1[1] + x2[1]; Iw[2] = Ia_n[2] + x1[2] + x2[2];
      Iw[3] = Ia_f[0] + x3[0]; Iw[4] = Ia_f[1] + x3[1]; Iw[5] = Ia_f[2] + x3[2];
    }
    // child -> parent accumulation of [f | composite inertia]: host-built schedule, every round each lane pulls
    // one finished child subtree (or nothing)
    {
      const int rounds = U.acc_rounds;
      for (int rd = 0; rd < rounds; rd++) {
        const int src = __ldg(&M->tsched[rd][lane]);
        const int s = src < 0 ? lane : src;
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const float o = SHW(Iw[k], s);
          if (src >= 0) Iw[k] += o;
        }
      }
    }
    if (is_dof) {
#pragma unroll
      for (int k = 0; k < 6; k++) sm.S[lane][k] = S[k];
    }
    // F = I^c S ; tau_bias = S . f^c
    float F[6];
    {
      const float ms = Iw[6];
      const float* h = &Iw[7];
      const float* I = &Iw[10];
      float t[3];
      cross3(h, S + 3, t);
      F[0] = I[0] * S[0] + I[1] * S[1] + I[2] * S[2] + t[0];
      F[1] = I[1] * S[0] + I[3] * S[1] + I[4] * S[2] + t[1];
      F[2] = I[2] * S[0] + I[4] * S[1] + I[5] * S[2] + t[2];
      cross3(h, S, t);
      F[3] = ms * S[3] - t[0]; F[4] = ms * S[4] - t[1]; F[5] = ms * S[5] - t[2];
    }
    const float tau_b = S[0] * Iw[0] + S[1] * Iw[1] + S[2] * Iw[2] + S[3] * Iw[3] + S[4] * Iw[4] + S[5] * Iw[5];
    // joint-space inertia: M[d][e] = S_e . F_d for every e on the path base..d; symmetric fill through smem
    for (int k = lane; k < 32 * 32; k += 32) sm.Mi[k] = 0.f;
    __syncwarp();
    {
      unsigned pm = is_dof ? __ldg(&M->link_dofmask[li]) : 0u;
      while (pm) {
        const int e = __ffs(pm) - 1;
        pm &= pm - 1;
        const float val = sm.S[e][0] * F[0] + sm.S[e][1] * F[1] + sm.S[e][2] * F[2] + sm.S[e][3] * F[3] + sm.S[e][4] * F[4] +
                          sm.S[e][5] * F[5];
        sm.Mi[MI(lane, e)] = val;
        sm.Mi[MI(e, lane)] = val;
      }
    }
    TREE_BARRIER();
    float a[32];
#pragma unroll
    for (int e = 0; e < 32; e++) a[e] = (is_dof && e < nd) ? sm.Mi[MI(lane, e)] : ((e == lane) ? 1.f : 0.f);
    // in-place Gauss-Jordan inverse (M is symmetric positive definite: no pivoting), lane = row.  The loop over the
    // pivots is ROLLED (the kernel is instruction-fetch bound when it is not): the row registers rotate one place per
    // step, so the pivot column is always a[0] — the FFMA that updates column j writes it to slot j-1, the finished
    // pivot column enters at slot 31, and after 32 steps every column is back in its own slot.
    // Body = 32 SHFL + 31 FFMA.
#pragma unroll 1
    for (int k = 0; k < 32; k++) {
      float rk[32];
#pragma unroll
      for (int j = 0; j < 32; j++) rk[j] = SHW(a[j], k);
      const float pinv = 1.0f / rk[0];
      const bool piv = lane == k;
      // the pivot lane holds the pivot row itself (a[j] == rk[j]): a[j] * pinv = a[j] + (pinv - 1) * rk[j], so one FFMA
      // form serves every lane
      const float coef = piv ? pinv - 1.0f : -a[0] * pinv;
#pragma unroll
      for (int j = 1; j < 32; j++) a[j - 1] = fmaf(coef, rk[j], a[j]);
      a[31] = piv ? pinv : coef;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; e++) sm.Mi[MI(lane, e)] = (is_dof && e < nd) ? a[e] : 0.f;
    // unconstrained acceleration and v*
    const float rhs_d = is_dof ? (-tau_b - __ldg(&M->joint_damping[li]) * my_qd) : 0.f;
    float qdd = 0.f;
#pragma unroll
    for (int e = 0; e < 32; e++) qdd = fmaf(a[e], SHW(rhs_d, e), qdd);
    const float vstar_d = is_dof ? my_qd + dt * qdd : 0.f;
    float cvs[3], cws[3];
    {
      const float vn = sqrtf(dot3(cv, cv)), wn = sqrtf(dot3(cw, cw));
#pragma unroll
      for (int k = 0; k < 3; k++) {
        cvs[k] = cv[k] + dt * (P.gravity[k] - cv[k] * (P.damp_lin_k1 + P.damp_lin_k2 * vn));
        cws[k] = cw[k] + dt * (-cw[k] * (P.damp_ang_k1 + P.damp_ang_k2 * wn));
      }
    }
    sm.vstar[lane] = vstar_d;
    if (lane < 8) sm.vstar[32 + lane] = lane < 3 ? cvs[lane] : (lane < 6 ? cws[lane - 3] : 0.f);
    TREE_BARRIER();

    // ---- collision detection (pre-step poses): same canonical order and keys as the Panda kernel ----
    TREE_PROBE(3);   // dynamics done
    float Rc[9];
    quat_to_mat(cquat, Rc);
    const float ca = P.cube_half, margin = P.contact_margin;
    bool v_hit = false;
    float v_pos[3] = {0, 0, 0}, v_dist = 0.f, v_top = 0.f, v_mu = 0.f;
    int v_key = 0;
    if (lane < 8) {
      const float l[3] = {(lane & 1) ? ca : -ca, (lane & 2) ? ca : -ca, (lane & 4) ? ca : -ca};
      m3vec(Rc, l, v_pos);
      v_pos[0] += cpos[0]; v_pos[1] += cpos[1]; v_pos[2] += cpos[2];
      const bool over = v_pos[0] >= P.table_min[0] && v_pos[0] <= P.table_max[0] && v_pos[1] >= P.table_min[1] &&
                        v_pos[1] <= P.table_max[1] && v_pos[2] > P.table_min[2];
      if (over) { v_top = P.table_max[2]; v_mu = P.cube_mu * P.table_mu; v_key = KEY_CUBE_TABLE + lane; }
      else { v_top = 0.f; v_mu = P.cube_mu * P.plane_mu; v_key = KEY_CUBE_PLANE + lane; }
      v_dist = v_pos[2] - v_top;
      v_hit = v_dist < margin;
    }
    const int ns = U.n_spheres;
    bool sc_hit = false, st_hit = false;
    float s_c[3] = {0, 0, 0}, sc_n[3] = {0, 0, 0}, sc_pB[3] = {0, 0, 0}, sc_dist = 0.f, st_dist = 0.f, s_r = 0.f;
    int s_link = 0;
    if (lane < ns) {
      s_link = __ldg(&M->sph_link[lane]);
      s_r = __ldg(&M->sph_r[lane]);
      const float lc[3] = {__ldg(&M->sph_c[lane][0]), __ldg(&M->sph_c[lane][1]), __ldg(&M->sph_c[lane][2])};
      float o[3], Rl[9];
#pragma unroll
      for (int k = 0; k < 9; k++) Rl[k] = sm.T[s_link][k];
      m3vec(Rl, lc, o);
      s_c[0] = sm.T[s_link][9] + o[0]; s_c[1] = sm.T[s_link][10] + o[1]; s_c[2] = sm.T[s_link][11] + o[2];
      const float rel[3] = {s_c[0] - cpos[0], s_c[1] - cpos[1], s_c[2] - cpos[2]};
      float l[3], cl[3];
      m3tvec(Rc, rel, l);
      bool inside = true;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        cl[j] = l[j] < -ca ? -ca : (l[j] > ca ? ca : l[j]);
        if (cl[j] != l[j]) inside = false;
      }
      float nloc[3];
      if (!inside) {
        const float dv[3] = {l[0] - cl[0], l[1] - cl[1], l[2] - cl[2]};
        const float d = sqrtf(dot3(dv, dv));
        sc_dist = d - s_r;
        sc_hit = sc_dist < margin;
        nloc[0] = dv[0] / d; nloc[1] = dv[1] / d; nloc[2] = dv[2] / d;
      } else {
        int axm = 0;
        float best = ca - fabsf(l[0]);
#pragma unroll
        for (int j = 1; j < 3; j++) {
          const float pen = ca - fabsf(l[j]);
          if (pen < best) { best = pen; axm = j; }
        }
        nloc[0] = nloc[1] = nloc[2] = 0.f;
        const float sgn = ((axm == 0 ? l[0] : (axm == 1 ? l[1] : l[2])) >= 0) ? 1.f : -1.f;
        if (axm == 0) { nloc[0] = sgn; cl[0] = sgn * ca; }
        else if (axm == 1) { nloc[1] = sgn; cl[1] = sgn * ca; }
        else { nloc[2] = sgn; cl[2] = sgn * ca; }
        sc_dist = -best - s_r;
        sc_hit = true;
      }
      if (sc_hit) {
        float pwl[3];
        m3vec(Rc, nloc, sc_n);
        m3vec(Rc, cl, pwl);
        sc_pB[0] = cpos[0] + pwl[0]; sc_pB[1] = cpos[1] + pwl[1]; sc_pB[2] = cpos[2] + pwl[2];
      }
      const bool over = s_c[0] >= P.table_min[0] && s_c[0] <= P.table_max[0] && s_c[1] >= P.table_min[1] &&
                        s_c[1] <= P.table_max[1] && s_c[2] > P.table_min[2];
      st_dist = s_c[2] - s_r - P.table_max[2];
      st_hit = over && (st_dist < margin);
    }
    // The static world beyond the table top (rim and sides of the slab, legs, ground plane) is the general path: an environment
    // takes it when its cube is not wholly over the top (bounding sphere, as the oracle decides) or when one of its spheres is
    // near anything but the top face (the oracle's own bounding culls).  Otherwise the closed forms above are exactly what
    // oracle collide() produces: vertices on the top plane, spheres over the top face.
    const unsigned lt = (1u << lane) - 1u;
    const float rb_c = ca * 1.7320508075688772f;
    const bool cube_fast = cpos[0] - rb_c >= P.table_min[0] && cpos[0] + rb_c <= P.table_max[0] && cpos[1] - rb_c >= P.table_min[1] &&
                           cpos[1] + rb_c <= P.table_max[1] && cpos[2] >= P.table_max[2];
    bool sph_general = false;
    if (lane < ns) {
      const int nsb = P.n_sboxes > 0 ? P.n_sboxes : 1;
      for (int k = 0; k < nsb; k++) {
        float bc[3], bh[3];
        sbox_get(P, k, bc, bh);
        const float d[3] = {s_c[0] - bc[0], s_c[1] - bc[1], s_c[2] - bc[2]};
        if (sqrtf(dot3(d, d)) - (s_r + sqrtf(dot3(bh, bh))) >= margin) continue;   // culled by the oracle as well
        const bool over_top = k == 0 && fabsf(d[0]) <= bh[0] && fabsf(d[1]) <= bh[1] && d[2] > bh[2];
        if (!over_top) sph_general = true;
      }
      if (s_c[2] - s_r < margin) sph_general = true;   // ground plane
    }
    if (!cube_fast || __any_sync(FULL, sph_general)) {
      const int cr = tree_collide_general(sm, M, P, lane, ns, maxc, cpos[0], cpos[1], cpos[2], cquat[0], cquat[1], cquat[2], cquat[3],
                                          cube_fast, v_pos[0], v_pos[1], v_pos[2], s_c[0], s_c[1], s_c[2], s_r, s_link, sc_hit, sc_dist,
                                          sc_n[0], sc_n[1], sc_n[2], sc_pB[0], sc_pB[1], sc_pB[2]);
      nc = cr & 255;
      if (cr & 256) flags |= B2E_ST_CONTACT_OVERFLOW;
    } else {
      const unsigned bv = __ballot_sync(FULL, v_hit), bsc = __ballot_sync(FULL, sc_hit), bst = __ballot_sync(FULL, st_hit);
      const int n_v = __popc(bv), n_sc = __popc(bsc), n_st = __popc(bst);
      const int total = n_v + n_sc + n_st;
      if (total > maxc) flags |= B2E_ST_CONTACT_OVERFLOW;
      nc = total > maxc ? maxc : total;
      if (v_hit) {
        const int cs = __popc(bv & lt);
        if (cs < maxc) {
          Contact& c = sm.con[cs];
          c.key = v_key; c.type = CT_CUBE_STATIC; c.link = -1;
          c.pA[0] = v_pos[0]; c.pA[1] = v_pos[1]; c.pA[2] = v_pos[2];
          c.pB[0] = v_pos[0]; c.pB[1] = v_pos[1]; c.pB[2] = v_top;
          c.n[0] = 0.f; c.n[1] = 0.f; c.n[2] = 1.f;
          c.dist = v_dist; c.mu = v_mu; c.erp = P.erp;
          sm.con_cfm[cs] = 0.f;
        }
      }
      if (sc_hit) {
        const int cs = n_v + __popc(bsc & lt);
        if (cs < maxc) {
          Contact& c = sm.con[cs];
          const float serp = __ldg(&M->sph_erp[lane]);
          c.key = KEY_SPHERE_CUBE + lane; c.type = CT_SPHERE_CUBE; c.link = s_link;
  #pragma unroll
          for (int j = 0; j < 3; j++) { c.n[j] = sc_n[j]; c.pB[j] = sc_pB[j]; c.pA[j] = s_c[j] - sc_n[j] * s_r; }
          c.dist = sc_dist; c.mu = P.cube_mu * __ldg(&M->sph_mu[lane]);
          c.erp = serp >= 0 ? serp : P.erp;
          sm.con_cfm[cs] = __ldg(&M->sph_cfm[lane]);
        }
      }
      if (st_hit) {
        const int cs = n_v + n_sc + __popc(bst & lt);
        if (cs < maxc) {
          Contact& c = sm.con[cs];
          const float serp = __ldg(&M->sph_erp[lane]);
          c.key = KEY_SPHERE_TABLE + lane; c.type = CT_SPHERE_STATIC; c.link = s_link;
          c.n[0] = 0.f; c.n[1] = 0.f; c.n[2] = 1.f;
          c.pA[0] = s_c[0]; c.pA[1] = s_c[1]; c.pA[2] = s_c[2] - s_r;
          c.pB[0] = s_c[0]; c.pB[1] = s_c[1]; c.pB[2] = P.table_max[2];
          c.dist = st_dist; c.mu = P.table_mu * __ldg(&M->sph_mu[lane]);
          c.erp = serp >= 0 ? serp : P.erp;
          sm.con_cfm[cs] = __ldg(&M->sph_cfm[lane]);
        }
      }
    }
    // joint-limit rows near a limit (lane = dof): order (dof, lower) then (dof, upper)
    const float dlo = my_q - my_lower, dup = my_upper - my_q;
    const float lmar = is_dof ? __ldg(&M->limit_margin[li]) : 0.f;
    const bool lo_hit = is_dof && dlo < lmar, up_hit = is_dof && dup < lmar;
    const unsigned blo = __ballot_sync(FULL, lo_hit), bup = __ballot_sync(FULL, up_hit);
    int nlim = __popc(blo) + __popc(bup);
    if (nlim > B2E_MAX_LIMROWS) { flags |= B2E_ST_LIMIT_OVERFLOW; nlim = B2E_MAX_LIMROWS; }
    if (lo_hit) {
      const int ls = __popc(blo & lt) + __popc(bup & lt);
      if (ls < B2E_MAX_LIMROWS) { sm.lim_d[ls] = lane; sm.lim_dist[ls] = dlo; }
    }
    if (up_hit) {
      const int ls = __popc(blo & lt) + __popc(bup & lt) + (lo_hit ? 1 : 0);
      if (ls < B2E_MAX_LIMROWS) { sm.lim_d[ls] = lane | (1 << 8); sm.lim_dist[ls] = dup; }
    }
    TREE_BARRIER();

    // ---- constraint rows: motor row of dof `lane`, generic row `lane` ----
    TREE_PROBE(4);   // collision done
    const int fric_start = nlim + nc;
    const int RG = nlim + 3 * nc;
    R = nd + RG;
    const float inv_dt = 1.0f / dt;
    const float cinv_m = 1.0f / P.cube_mass, cinv_I = 1.0f / P.cube_inertia;
    MotorRegs m;
    {
      float desired = my_kp * (my_target - my_q) * inv_dt;
      const float mv = ik_cap ? P.ik_max_vel : __ldg(&M->max_vel[li]);
      if (mv > 0) desired = fminf(fmaxf(desired, -mv), mv);
      const float diag = is_dof ? sm.Mi[MI(lane, lane)] : 1.f;
      m.diag = diag;
      m.invd = is_dof ? 1.0f / diag : 0.f;
      m.u = is_dof ? desired - vstar_d : 0.f;
      m.hi = is_dof ? __ldg(&M->max_force[li]) * dt : 0.f;
      m.lo = -m.hi;
      m.lam = 0.f;
      m.prev = 0.f;
    }
    TreeRow rr;
    bool coupled, has_cube, arm_part, arm_contact;
    float Jc[6];   // cube part of this row's Jacobian
    {
      const int gi = lane;
      const bool valid = gi < RG;
      float J[32];
#pragma unroll
      for (int k = 0; k < 32; k++) J[k] = 0.f;
#pragma unroll
      for (int k = 0; k < 6; k++) Jc[k] = 0.f;
      float desired = 0.f, cfm = 0.f, lo = 0.f, hi = 0.f, mu = 0.f;
      int type = ROW_LIMIT, isl = 0, nidx = 0;
      bool sphere_cube_normal = false;
      const bool is_lim = valid && gi < nlim;
      int lim_dof = 0;
      float lim_sg = 0.f;
      if (valid) {
        if (gi < nlim) {
          const int code = sm.lim_d[gi];
          const int d = code & 0xff, side = code >> 8;
          lim_dof = d;
          lim_sg = side ? -1.f : 1.f;
#pragma unroll
          for (int k = 0; k < 32; k++) J[k] = (k == d) ? (side ? -1.f : 1.f) : 0.f;
          const float pen = sm.lim_dist[gi] + P.slop;
          desired = pen > 0 ? -pen * inv_dt : -pen * P.erp * inv_dt;
          lo = 0.f; hi = 1e30f;
        } else {
          int c, which;
          if (gi < fric_start) { c = gi - nlim; which = 0; }
          else { c = (gi - fric_start) >> 1; which = 1 + ((gi - fric_start) & 1); }
          const Contact& ct = sm.con[c];
          const float n[3] = {ct.n[0], ct.n[1], ct.n[2]};
          float dir[3];
          if (which == 0) { dir[0] = n[0]; dir[1] = n[1]; dir[2] = n[2]; }
          else {
            float t1[3], t2[3];
            plane_space(n, t1, t2);
            if (which == 1) { dir[0] = t1[0]; dir[1] = t1[1]; dir[2] = t1[2]; }
            else { dir[0] = t2[0]; dir[1] = t2[1]; dir[2] = t2[2]; }
          }
          isl = ct.type == CT_CUBE_STATIC ? 1 : 0;
          if (ct.type == CT_CUBE_STATIC) {
            const float rel[3] = {ct.pA[0] - cpos[0], ct.pA[1] - cpos[1], ct.pA[2] - cpos[2]};
            float t[3];
            cross3(rel, dir, t);
#pragma unroll
            for (int k = 0; k < 3; k++) { Jc[k] = dir[k]; Jc[3 + k] = t[k]; }
          } else {
            const unsigned mask = __ldg(&M->link_dofmask[ct.link]);
            const float rel[3] = {ct.pA[0] - U.base_pos[0], ct.pA[1] - U.base_pos[1], ct.pA[2] - U.base_pos[2]};
            float wn[3];
            cross3(rel, dir, wn);
#pragma unroll
            for (int d = 0; d < 32; d++) {
              const float vv = sm.S[d][0] * wn[0] + sm.S[d][1] * wn[1] + sm.S[d][2] * wn[2] + sm.S[d][3] * dir[0] +
                               sm.S[d][4] * dir[1] + sm.S[d][5] * dir[2];
              J[d] = ((mask >> d) & 1u) ? vv : 0.f;
            }
            if (ct.type == CT_SPHERE_CUBE) {
              const float relc[3] = {ct.pB[0] - cpos[0], ct.pB[1] - cpos[1], ct.pB[2] - cpos[2]};
              float t[3];
              cross3(relc, dir, t);
#pragma unroll
              for (int k = 0; k < 3; k++) { Jc[k] = -dir[k]; Jc[3 + k] = -t[k]; }
              sphere_cube_normal = (which == 0);
            }
          }
          if (which == 0) {
            type = ROW_NORMAL;
            const float pen = ct.dist + P.slop;
            desired = pen > 0 ? -pen * inv_dt : -pen * ct.erp * inv_dt;
            lo = 0.f; hi = 1e30f;
            cfm = sm.con_cfm[c];
          } else {
            type = ROW_FRICTION;
            mu = ct.mu;
            nidx = nlim + c;
          }
        }
      }
      arm_part = valid && isl == 0;
      // W = M^-1 J^T (rows of cube-table contacts have no arm part)
      // Without a proxy contact (the common case) the only rows with an arm part are the limit rows, J = -+e_d:
      // their W row is a signed column of M^-1 and every J . x below is one signed element (the same bits as the
      // dense sums, whose other terms are exact zeros).
      float diag = 0.f, jv = 0.f;
      arm_contact = __any_sync(FULL, arm_part && gi >= nlim);
      if (arm_contact) {
#pragma unroll 4
        for (int d = 0; d < 32; d++) {
          float acc = 0.f;
#pragma unroll
          for (int e = 0; e < 32; e++) acc = fmaf(sm.Mi[MI(d, e)], J[e], acc);
          sm.W[gi * TREE_WS + d] = acc;
        }
        __syncwarp();
#pragma unroll
        for (int d = 0; d < 32; d++) {
          diag = fmaf(J[d], sm.W[gi * TREE_WS + d], diag);
          jv = fmaf(J[d], sm.vstar[d], jv);
        }
      } else {
#pragma unroll 4
        for (int d = 0; d < 32; d++) sm.W[gi * TREE_WS + d] = is_lim ? lim_sg * sm.Mi[MI(d, lim_dof)] : 0.f;
        __syncwarp();
        if (is_lim) {
          diag = lim_sg * sm.W[gi * TREE_WS + lim_dof];
          jv = lim_sg * sm.vstar[lim_dof];
        }
      }
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const float wv = Jc[k] * (k < 3 ? cinv_m : cinv_I);
        sm.W[gi * TREE_WS + 32 + k] = wv;
        diag = fmaf(Jc[k], wv, diag);
        jv = fmaf(Jc[k], sm.vstar[32 + k], jv);
      }
      sm.W[gi * TREE_WS + 38] = 0.f; sm.W[gi * TREE_WS + 39] = 0.f; sm.W[gi * TREE_WS + 40] = 0.f;
      rr.type = type; rr.isl = isl; rr.nidx = nidx;
      rr.lo = lo; rr.hi = hi; rr.mu = mu;
      rr.diag = valid ? diag + cfm : 0.f;
      rr.invd = valid ? 1.0f / (diag + cfm) : 0.f;
      rr.gg = 1.0f - cfm * rr.invd;
      rr.u = valid ? desired - jv : 0.f;
      rr.lam = 0.f; rr.base = 0.f; rr.prev = 0.f;
      coupled = __any_sync(FULL, sphere_cube_normal);
      has_cube = __any_sync(FULL, valid && isl == 1);
      __syncwarp();
      // generic x generic block: A[c][r] = J_r . W_c
      for (int c = 0; c < RG; c++) {
        float acc = 0.f;
        if (arm_contact) {
#pragma unroll
          for (int k = 0; k < 32; k++) acc = fmaf(J[k], sm.W[c * TREE_WS + k], acc);
        } else if (is_lim) {
          acc = lim_sg * sm.W[c * TREE_WS + lim_dof];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) acc = fmaf(Jc[k], sm.W[c * TREE_WS + 32 + k], acc);
        sm.A[c * 32 + gi] = valid ? acc : 0.f;
      }
      __syncwarp();
      // warm start (contact rows): lambda0 = cached impulse * factor, u -= A lambda0
      if (gi >= nlim && gi < RG) {
        int c, j;
        if (gi < fric_start) { c = gi - nlim; j = 0; }
        else { c = (gi - fric_start) >> 1; j = 1 + ((gi - fric_start) & 1); }
        const int key = sm.con[c].key;
        float l0 = 0.f;
        for (int sl = 0; sl < B2E_CACHE_SLOTS; sl++)
          if (sm.ckey[sl] == key) { l0 = sm.clam[sl][j] * P.warmstart; break; }
        rr.lam = l0;
        rr.base = l0 * rr.gg;
      }
      for (int c = nlim; c < RG; c++) {
        const float l0 = SHW(rr.lam, c);
        if (l0 != 0.f) {
          rr.u = fmaf(-sm.A[c * 32 + lane], l0, rr.u);
          m.u = fmaf(-sm.W[c * TREE_WS + lane], l0, m.u);
        }
      }
    }
    TREE_BARRIER();
    bool arm_done = false;
    int iters_arm = 0;
    TREE_PROBE(5);   // rows, Delassus block, warm start done
    {
      // arm island = motor rows + limit rows only (no proxy contact): affine Gauss-Seidel
      if (!coupled && !arm_contact) {
        float lam_m = 0.f, lam_l[3];
        const float r0 = SHW(rr.u, 0), r1 = SHW(rr.u, 1), r2 = SHW(rr.u, 2);
        const int ia = tree_arm_affine(sm, lane, nd, nlim, m.u, m.invd, m.diag, m.lo, m.hi, r0, r1, r2, P.solver_iters,
                                       P.residual_tol, lam_m, lam_l);
        if (ia >= 0) {
          arm_done = true;
          iters_arm = ia;
          m.lam = lam_m;
          if (lane < nlim) rr.lam = lane == 0 ? lam_l[0] : (lane == 1 ? lam_l[1] : lam_l[2]);
        }
      }
    }
    TREE_PROBE(6);   // affine arm island done
    iters = tree_pgs(sm, lane, m, rr, nd, RG, fric_start, coupled, has_cube, arm_done, P.solver_iters, P.residual_tol);
    TREE_PROBE(7);   // sweeps done
    if (iters_arm > iters) iters = iters_arm;
    sm.mlam[lane] = is_dof ? m.lam : 0.f;
    sm.glam[lane] = lane < RG ? rr.lam : 0.f;
#if TREE_BAR_POST_SOLVE
    TREE_BARRIER();
#else
    __syncwarp();
#endif

    // ---- delta velocities dv = sum_r W_r lambda_r ----
    TREE_PROBE(8);
    float dvk = 0.f, dvc = 0.f;
    for (int g2 = 0; g2 < RG; g2++) {
      const float lg = sm.glam[g2];
      dvk = fmaf(sm.W[g2 * TREE_WS + lane], lg, dvk);
      if (lane < 6) dvc = fmaf(sm.W[g2 * TREE_WS + 32 + lane], lg, dvc);
    }
#pragma unroll 8
    for (int d = 0; d < 32; d++) dvk = fmaf(sm.Mi[MI(lane, d)], sm.mlam[d], dvk);
    // ---- integrate (semi-implicit Euler) ----
    if (is_dof && !ghost) {
      my_qd = vstar_d + dvk;
      my_q += dt * my_qd;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float dvl = SHW(dvc, k), dva = SHW(dvc, 3 + k);
      if (!ghost) {
        cv[k] = cvs[k] + dvl;
        cw[k] = cws[k] + dva;
        cpos[k] += dt * cv[k];
      }
    }
    if (!ghost) {
      const float wl = sqrtf(dot3(cw, cw)), ang = wl * dt;
      float f, cs;
      if (ang < 1e-3f) { f = 0.5f * dt - dt * dt * dt * (1.0f / 48.0f) * wl * wl; cs = cosf(0.5f * ang); }
      else { float sn; sincosf(0.5f * ang, &sn, &cs); f = sn / wl; }
      const float dq[4] = {cw[0] * f, cw[1] * f, cw[2] * f, cs};
      float nq[4];
      quat_mul(dq, cquat, nq);
      const float nn = 1.0f / sqrtf(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
#pragma unroll
      for (int k = 0; k < 4; k++) cquat[k] = nq[k] * nn;
    }
    // ---- contact cache for the next step's warm start ----
    TREE_PROBE(9);   // integrated
    {
      int key = -1;
      float l3[3] = {0.f, 0.f, 0.f};
      if (lane < nc) {
        key = sm.con[lane].key;
        l3[0] = sm.glam[nlim + lane];
        l3[1] = sm.glam[nlim + nc + 2 * lane];
        l3[2] = sm.glam[nlim + nc + 2 * lane + 1];
      }
      if (record_contacts && lane < B2E_MAX_CONTACTS && live_env && !ghost) {
        float* o = st.contacts + ((size_t)env * B2E_MAX_CONTACTS + lane) * 8;
        if (lane < nc) {
          o[0] = (float)key; o[1] = sm.con[lane].dist;
          o[2] = sm.con[lane].n[0]; o[3] = sm.con[lane].n[1]; o[4] = sm.con[lane].n[2];
          o[5] = l3[0]; o[6] = l3[1]; o[7] = l3[2];
        } else {
#pragma unroll
          for (int k = 0; k < 8; k++) o[k] = 0.f;
        }
      }
      __syncwarp();
      if (lane < B2E_CACHE_SLOTS && !ghost) {
        sm.ckey[lane] = key;
        sm.clam[lane][0] = l3[0]; sm.clam[lane][1] = l3[1]; sm.clam[lane][2] = l3[2];
      }
      __syncwarp();
    }
    {
      bool bad = is_dof && !(isfinite(my_q) && isfinite(my_qd));
      bad = bad || !(isfinite(cpos[0]) && isfinite(cpos[1]) && isfinite(cpos[2]) && isfinite(cv[0]) && isfinite(cv[1]) &&
                     isfinite(cv[2]) && isfinite(cw[0]) && isfinite(cw[1]) && isfinite(cw[2]));
      if (__any_sync(FULL, bad)) flags |= B2E_ST_NAN;
    }
  }

  // ---- store state ----
  TREE_PROBE(10);
  if (is_dof && live_env) {
    st.q[env * nd + lane] = my_q;
    st.qd[env * nd + lane] = my_qd;
    st.mtarget[env * nd + lane] = my_target;
  }
  if (lane == 0 && live_env) {
    st.obj_pose[env * 7 + 0] = cpos[0]; st.obj_pose[env * 7 + 1] = cpos[1]; st.obj_pose[env * 7 + 2] = cpos[2];
    st.obj_pose[env * 7 + 3] = cquat[0]; st.obj_pose[env * 7 + 4] = cquat[1];
    st.obj_pose[env * 7 + 5] = cquat[2]; st.obj_pose[env * 7 + 6] = cquat[3];
    st.obj_vel[env * 6 + 0] = cv[0]; st.obj_vel[env * 6 + 1] = cv[1]; st.obj_vel[env * 6 + 2] = cv[2];
    st.obj_vel[env * 6 + 3] = cw[0]; st.obj_vel[env * 6 + 4] = cw[1]; st.obj_vel[env * 6 + 5] = cw[2];
  }
  if (IK && lane < 6 && live_env) st.hand_pose[env * 6 + lane] = my_hp;
  if (lane < B2E_CACHE_SLOTS && live_env) {
    st.cache_key[env * B2E_CACHE_SLOTS + lane] = sm.ckey[lane];
#pragma unroll
    for (int j = 0; j < 3; j++) st.cache_lam[(env * B2E_CACHE_SLOTS + lane) * 3 + j] = sm.clam[lane][j];
  }

  // ---- observation / termination / reward (icub_push_gym_env.py:276-282) ----
  TREE_PROBE(11);
  if (mode == B2E_MODE_ACTION || obs_out) {
    const int ee = U.ee_link;
    const float cm[3] = {U.ee_com[0], U.ee_com[1], U.ee_com[2]};
    float o[3];
    m3vec(Rm, cm, o);
    float epos[3], Re[9];
#pragma unroll
    for (int k = 0; k < 3; k++) epos[k] = SHW(pw[k] + o[k], ee);
#pragma unroll
    for (int k = 0; k < 9; k++) Re[k] = SHW(Rm[k], ee);
    // linear velocity of the hand COM: sum over the joints on the path of (axis x (p_ee - o_j)) qd_j
    float vl[3] = {0, 0, 0};
    {
      const int jt = __ldg(&M->jtype[li]);
      const float ax[3] = {__ldg(&M->axis[li][0]), __ldg(&M->axis[li][1]), __ldg(&M->axis[li][2])};
      float aw[3];
      m3vec(Rm, ax, aw);
      if (is_dof && ((U.ee_dofmask >> lane) & 1u)) {
        if (jt == B2E_JOINT_REVOLUTE) {
          const float rel[3] = {epos[0] - pw[0], epos[1] - pw[1], epos[2] - pw[2]};
          float t[3];
          cross3(aw, rel, t);
          vl[0] = t[0] * my_qd; vl[1] = t[1] * my_qd; vl[2] = t[2] * my_qd;
        } else {
          vl[0] = aw[0] * my_qd; vl[1] = aw[1] * my_qd; vl[2] = aw[2] * my_qd;
        }
      }
      // fixed-order sum over the bodies (the oracle's order)
      float acc[3] = {0, 0, 0};
      for (int j = 0; j < nl; j++) {
        acc[0] += SHW(vl[0], j); acc[1] += SHW(vl[1], j); acc[2] += SHW(vl[2], j);
      }
      vl[0] = acc[0]; vl[1] = acc[1]; vl[2] = acc[2];
    }
    float equat[4], eeu[3], ceu[3];
    mat_to_quat(Re, equat);
    quat_to_euler(equat, eeu);
    quat_to_euler(cquat, ceu);
    float hq[4], oq[4], hqi[4], relq[4], releu[3], relp[3];
    euler_to_quat(eeu, hq);
    euler_to_quat(ceu, oq);
    hqi[0] = -hq[0]; hqi[1] = -hq[1]; hqi[2] = -hq[2]; hqi[3] = hq[3];
    {
      const float dd[3] = {cpos[0] - epos[0], cpos[1] - epos[1], cpos[2] - epos[2]};
      float Rh[9];
      quat_to_mat(hqi, Rh);
      m3vec(Rh, dd, relp);
    }
    quat_mul(hqi, oq, relq);
    quat_to_euler(relq, releu);
    float* obsb = sm.scr;
    __syncwarp();
    if (lane == 0) {
      int n = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = epos[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = eeu[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = (vl[k] - P.vel_mean[k]) / P.vel_std[k];
      n += n_qobs;
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = cpos[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = ceu[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = relp[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = releu[k];
      if (P.task == B2E_TASK_PUSH) {
#pragma unroll
        for (int k = 0; k < 3; k++) obsb[n++] = target[k];
      }
    }
    if (obs_idx >= 0) obsb[9 + obs_idx] = my_q;
    __syncwarp();
    for (int k = lane; k < P.n_obs; k += 32) {
      const float raw = obsb[k];
      if (!live_env) continue;
      st.raw_obs[(size_t)env * P.n_obs + k] = raw;
      if (obs_out) obs_out[(size_t)env * P.n_obs + k] = 2.0f * ((raw - P.obs_low[k]) / (P.obs_high[k] - P.obs_low[k])) - 1.0f;
    }
    const float dd1[3] = {epos[0] - cpos[0], epos[1] - cpos[1], epos[2] - cpos[2]};
    const float d1 = sqrtf(dot3(dd1, dd1));
    const float dd2[3] = {cpos[0] - target[0], cpos[1] - target[1], cpos[2] - target[2]};
    const float d2 = sqrtf(dot3(dd2, dd2));
    float rew;
    int dn = 0;
    if (P.task == B2E_TASK_PUSH && P.goal_env) {   // GoalEnv variant (icub_push_gym_goal_env.py:99-132)
      dn = (counter > P.max_steps) || (d2 <= P.dist_min);
      rew = d2 > P.dist_min ? -1.0f : 0.0f;
    } else if (P.reward_kind == B2E_REWARD_ICUB_REACH) {   // icub_reach_gym_env.py:301-330: the bonus is ADDED
      if (d1 <= P.dist_min) { terminated = 1; dn = 1; }
      else if (terminated || counter > P.max_steps) dn = 1;
      rew = -d1;
      if (d1 <= P.dist_min) rew += 1000.0f + (100.0f - d1 * 80.0f);
    } else if (P.reward_kind == B2E_REWARD_ICUB_PUSH0 || P.reward_kind == B2E_REWARD_ICUB_PUSH1) {   // icub_push_gym_env.py:327-373
      if (d2 <= P.dist_min) { terminated = 1; dn = 1; }
      else if (terminated || counter > P.max_steps) dn = 1;
      if (P.reward_kind == B2E_REWARD_ICUB_PUSH0) rew = -d1 - d2;
      else {
        const float d0 = st.shaping[env * 2], dmax = st.shaping[env * 2 + 1];
        rew = 0.125f * (1.0f - d1 / d0);
        if (!(d1 > 0.1f)) rew += 0.25f * (1.0f - d2 / dmax);
      }
      if (d2 <= P.dist_min) rew += 1000.0f;
    } else if (P.task == B2E_TASK_PUSH) {
      if (d2 <= P.dist_min) { terminated = 1; dn = 1; }
      else if (terminated || counter > P.max_steps) dn = 1;
      rew = -d1 - d2;
      if (d2 <= P.dist_min) rew = 1000.0f + (100.0f - d2 * 80.0f);
    } else {
      if (d1 <= P.dist_min) { terminated = 1; dn = 1; }
      else if (terminated || counter > P.max_steps) dn = 1;
      rew = -d1;
      if (d1 <= P.dist_min) rew = 1000.0f + (100.0f - d1 * 80.0f);
    }
    if (lane == 0 && live_env) {
      if (reward_out) reward_out[env] = rew;
      if (done_out) done_out[env] = (float)dn;
    }
  }
  if (lane == 0 && mode != B2E_MODE_OBSERVE && live_env) {
    st.counters[env * 2] = counter;
    st.counters[env * 2 + 1] = terminated;
    st.status[env * 4 + 0] = flags;
    if (nsub > 0) {
      st.status[env * 4 + 1] = iters;
      st.status[env * 4 + 2] = nc;
      st.status[env * 4 + 3] = R;
    }
  }
  TREE_PROBE(12);   // observation / reward done
#ifdef PROFILE_WARM
  if (lane == 0 && live_env && nsub > 0)
    for (int k = 0; k < 16; k++) st.contacts[(size_t)env * B2E_MAX_CONTACTS * 8 + 32 + k] = (float)tp_am[k];
#endif
  if (sched && lane == 0 && live_env) {   // file this environment under its cost class for the next full-batch step
    const int cls = min(NBK_MAIN - 1, iters / 10);
    const int pos = atomicAdd(&sched[SCHED_CNT(seq & 3, cls)], 1);
    sched[SCHED_LIST(seq & 1, cls, st.B) + pos] = env;
  }
}
