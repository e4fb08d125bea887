/*
 * b2oracle.c — CPU restatement of the reference's env.step() hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libb2env.so) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference path lives in the third-party `pybullet`
 * wheel (Bullet3 C++; un-pinned in the reference's requirements.txt:2, `pybullet==2.5.0` hinted at
 * setup.py:35), which is not vendored in /root/reference and not installable here, and the
 * reference ships no tests or golden vectors.  This file therefore restates
 * (a) the Python wrapper semantics, which ARE in the reference and are followed line by line, and
 * (b) the published structure of Bullet's btMultiBodyDynamicsWorld step as recalled in SURVEY.md
 * Appendix D — Featherstone ABA, velocity-level motor/limit/contact rows, projected Gauss-Seidel
 * in delta-velocity space with M^-1 J^T obtained by an O(n) ABA-style pass.  It is pinned by the
 * first-principles known-answer vectors of SURVEY.md Appendix C (tests/test_oracle_kat.py) and by
 * closed-form checks, not by PyBullet output.
 *
 * Reference call sites restated (paths relative to pybullet_robot_envs/envs/):
 *   step()                      panda_envs/panda_push_gym_env.py:244-255, panda_reach_gym_env.py:228-239
 *   apply_action() (task)       panda_push_gym_env.py:189-242
 *   apply_action() (robot)      panda_envs/panda_env.py:293-310 (joint mode), :229-291 (IK mode)
 *   p.stepSimulation            panda_push_gym_env.py:236       [Bullet, EXT-recalled]
 *   get_observation() (robot)   panda_env.py:141-193
 *   get_observation() (world)   world_envs/world_env.py:109-126
 *   get_extended_observation()  panda_push_gym_env.py:150-187
 *   scale_gym_data()            utils.py:78-91
 *   _termination()              panda_push_gym_env.py:301-316, panda_reach_gym_env.py:285-301
 *   _compute_reward()           panda_push_gym_env.py:318-331, panda_reach_gym_env.py:303-313
 *
 * Build: see oracle/Makefile.  `real` is float by default (to compare against the fp32 GPU
 * path); -DB2O_DOUBLE computes in double (used to bound the fp32 error).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b2env.h"

#ifdef B2O_DOUBLE
typedef double real;
#define RSQRT sqrt
#define RSIN sin
#define RCOS cos
#define RATAN2 atan2
#define RASIN asin
#define RFABS fabs
#else
typedef float real;
#define RSQRT sqrtf
#define RSIN sinf
#define RCOS cosf
#define RATAN2 atan2f
#define RASIN asinf
#define RFABS fabsf
#endif

#define NL B2E_MAX_LINKS
#define ND B2E_MAX_DOF
#define NV (B2E_MAX_DOF + 6)
#define MAXC B2E_MAX_CONTACTS
#define MAXROWS (B2E_MAX_DOF + B2E_MAX_LIMROWS + 3 * B2E_MAX_CONTACTS)

typedef struct b2o_state {
  int32_t B;
  float* q;          /* [B][n_dof] */
  float* qd;         /* [B][n_dof] */
  float* obj_pose;   /* [B][7] */
  float* obj_vel;    /* [B][6] */
  float* target;     /* [B][3] */
  float* mtarget;    /* [B][n_dof] */
  int32_t* counters; /* [B][2] */
  int32_t* cache_key;/* [B][16] */
  float* cache_lam;  /* [B][16][3] */
  float* hand_pose;  /* [B][6] */
  int32_t* status;   /* [B][4] */
  float* raw_obs;    /* [B][n_obs] */
  float* contacts;   /* [B][12][8] */
  float* shaping;    /* [B][2] */
} b2o_state;

/* Joint lists of the task (b2e_params.n_obs_joints > 0: iCub, icub_env.py:121-143; else Panda: identity). */
static int is_ctrl_dof(const b2e_params* P, int d) {
  return P->n_obs_joints > 0 ? (int)((P->ctrl_mask >> d) & 1u) : d < P->n_ctrl;
}
static int ctrl_dof_of(const b2e_params* P, int k) { return P->n_obs_joints > 0 ? P->ctrl_dof[k] : k; }

/* ------------------------------------------------------------------ small math */
static void m3_mul(const real* a, const real* b, real* c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
static void m3_vec(const real* a, const real* v, real* o) {
  real x = a[0] * v[0] + a[1] * v[1] + a[2] * v[2];
  real y = a[3] * v[0] + a[4] * v[1] + a[5] * v[2];
  real z = a[6] * v[0] + a[7] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void m3t_vec(const real* a, const real* v, real* o) {
  real x = a[0] * v[0] + a[3] * v[1] + a[6] * v[2];
  real y = a[1] * v[0] + a[4] * v[1] + a[7] * v[2];
  real z = a[2] * v[0] + a[5] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void cross3(const real* a, const real* b, real* o) {
  real x = a[1] * b[2] - a[2] * b[1];
  real y = a[2] * b[0] - a[0] * b[2];
  real z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* Rodrigues rotation about unit axis by angle */
static void axis_angle(const real* ax, real ang, real* R) {
  real c = RCOS(ang), s = RSIN(ang), t = 1 - c;
  real x = ax[0], y = ax[1], z = ax[2];
  R[0] = t * x * x + c;     R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
  R[3] = t * x * y + s * z; R[4] = t * y * y + c;     R[5] = t * y * z - s * x;
  R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}

/* quaternion xyzw */
static void quat_to_mat(const real* q, real* R) {
  real x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}
static void mat_to_quat(const real* R, real* q) {
  real tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    real s = RSQRT(tr + 1) * 2;
    q[3] = s / 4; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    real s = RSQRT(1 + R[0] - R[4] - R[8]) * 2;
    q[3] = (R[7] - R[5]) / s; q[0] = s / 4; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    real s = RSQRT(1 + R[4] - R[0] - R[8]) * 2;
    q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = s / 4; q[2] = (R[5] + R[7]) / s;
  } else {
    real s = RSQRT(1 + R[8] - R[0] - R[4]) * 2;
    q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = s / 4;
  }
}
static void quat_mul(const real* a, const real* b, real* o) { /* o = a*b */
  real x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  real y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  real z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  real w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void quat_rot(const real* q, const real* v, real* o) {
  real R[9];
  quat_to_mat(q, R);
  m3_vec(R, v, o);
}
/* p.getEulerFromQuaternion (btQuaternion::getEulerZYX) [EXT-recalled]; used at panda_env.py:160,
 * world_env.py:119, panda_push_gym_env.py:176 */
static void quat_to_euler(const real* q, real* e) {
  real x = q[0], y = q[1], z = q[2], w = q[3];
  real sqx = x * x, sqy = y * y, sqz = z * z, sqw = w * w;
  real sarg = -2 * (x * z - w * y);
  if (sarg <= (real)-0.99999) {
    e[0] = 0; e[1] = (real)(-0.5 * M_PI); e[2] = 2 * RATAN2(x, -y);
  } else if (sarg >= (real)0.99999) {
    e[0] = 0; e[1] = (real)(0.5 * M_PI); e[2] = 2 * RATAN2(-x, y);
  } else {
    e[0] = RATAN2(2 * (y * z + w * x), sqw - sqx - sqy + sqz);
    e[1] = RASIN(sarg);
    e[2] = RATAN2(2 * (x * y + w * z), sqw + sqx - sqy - sqz);
  }
}
/* p.getQuaternionFromEuler (btQuaternion::setEulerZYX) [EXT-recalled]; panda_push_gym_env.py:169,172 */
static void euler_to_quat(const real* e, real* q) {
  real hr = e[0] * (real)0.5, hp = e[1] * (real)0.5, hy = e[2] * (real)0.5;
  real cr = RCOS(hr), sr = RSIN(hr), cp = RCOS(hp), sp = RSIN(hp), cy = RCOS(hy), sy = RSIN(hy);
  q[0] = sr * cp * cy - cr * sp * sy;
  q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy;
  q[3] = cr * cp * cy + sr * sp * sy;
}

/* ------------------------------------------------------------------ kinematics */
typedef struct {
  real R[NL][9]; /* world <- link rotation */
  real p[NL][3]; /* link origin in world    */
  real E[NL][9]; /* parent <- link rotation (local) */
  real r[NL][3]; /* link origin in parent frame      */
} fk_t;

static void joint_local(const b2e_model* m, int i, real qi, real* Rl, real* rl) {
  real jr[9], ax[3];
  for (int k = 0; k < 9; k++) jr[k] = m->jrot[i][k];
  for (int k = 0; k < 3; k++) { ax[k] = m->axis[i][k]; rl[k] = m->jpos[i][k]; }
  if (m->jtype[i] == B2E_JOINT_REVOLUTE) {
    real Rq[9];
    axis_angle(ax, qi, Rq);
    m3_mul(jr, Rq, Rl);
  } else if (m->jtype[i] == B2E_JOINT_PRISMATIC) {
    real t[3] = {ax[0] * qi, ax[1] * qi, ax[2] * qi}, o[3];
    m3_vec(jr, t, o);
    for (int k = 0; k < 3; k++) { rl[k] += o[k]; }
    memcpy(Rl, jr, sizeof(jr));
  } else {
    memcpy(Rl, jr, sizeof(jr));
  }
}

static void forward_kinematics(const b2e_model* m, const real* q, fk_t* fk) {
  real Rb[9], pb[3];
  for (int k = 0; k < 9; k++) Rb[k] = m->base_rot[k];
  for (int k = 0; k < 3; k++) pb[k] = m->base_pos[k];
  for (int i = 0; i < m->n_links; i++) {
    real qi = m->dof[i] >= 0 ? q[m->dof[i]] : 0;
    joint_local(m, i, qi, fk->E[i], fk->r[i]);
    const real* Rp = m->parent[i] < 0 ? Rb : fk->R[m->parent[i]];
    const real* pp = m->parent[i] < 0 ? pb : fk->p[m->parent[i]];
    m3_mul(Rp, fk->E[i], fk->R[i]);
    real o[3];
    m3_vec(Rp, fk->r[i], o);
    for (int k = 0; k < 3; k++) fk->p[i][k] = pp[k] + o[k];
  }
}

static int is_ancestor_or_self(const b2e_model* m, int a, int link) {
  while (link >= 0) {
    if (link == a) return 1;
    link = m->parent[link];
  }
  return 0;
}

/* world-frame point Jacobian: velocity of world point `pt` rigidly attached to `link`.
 * Jl[d] (3) linear, Ja[d] (3) angular, for every dof d */
static void point_jacobian(const b2e_model* m, const fk_t* fk, int link, const real* pt, real Jl[][3], real Ja[][3]) {
  for (int d = 0; d < m->n_dof; d++)
    for (int k = 0; k < 3; k++) { Jl[d][k] = 0; Ja[d][k] = 0; }
  for (int j = 0; j < m->n_links; j++) {
    int d = m->dof[j];
    if (d < 0 || !is_ancestor_or_self(m, j, link)) continue;
    real ax[3] = {m->axis[j][0], m->axis[j][1], m->axis[j][2]}, aw[3];
    m3_vec(fk->R[j], ax, aw);
    if (m->jtype[j] == B2E_JOINT_REVOLUTE) {
      real rel[3] = {pt[0] - fk->p[j][0], pt[1] - fk->p[j][1], pt[2] - fk->p[j][2]};
      cross3(aw, rel, Jl[d]);
      for (int k = 0; k < 3; k++) Ja[d][k] = aw[k];
    } else {
      for (int k = 0; k < 3; k++) Jl[d][k] = aw[k];
    }
  }
}

/* ------------------------------------------------------------------ spatial algebra (link coordinates)
 * motion vector [w; v], force vector [n; f]; transform parent->child given E (parent<-child) and r. */
static void xform_motion(const real* E, const real* r, const real* vp, real* vc) {
  /* w_c = E^T w_p ; v_c = E^T (v_p - r x w_p) */
  real t[3], u[3];
  cross3(r, vp, t);
  for (int k = 0; k < 3; k++) u[k] = vp[3 + k] - t[k];
  m3t_vec(E, vp, vc);
  m3t_vec(E, u, vc + 3);
}
static void xform_force_T(const real* E, const real* r, const real* fc, real* fp) {
  /* child -> parent: f_p = E f_c ; n_p = E n_c + r x (E f_c) */
  real f[3], n[3], t[3];
  m3_vec(E, fc + 3, f);
  m3_vec(E, fc, n);
  cross3(r, f, t);
  for (int k = 0; k < 3; k++) { fp[k] = n[k] + t[k]; fp[3 + k] = f[k]; }
}
static void x6(const real* E, const real* r, real X[36]) {
  /* 6x6 motion transform parent->child: [Et 0; -Et rx, Et] */
  real Et[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Et[3 * i + j] = E[3 * j + i];
  real rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
  real Etrx[9];
  m3_mul(Et, rx, Etrx);
  memset(X, 0, 36 * sizeof(real));
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      X[6 * i + j] = Et[3 * i + j];
      X[6 * (i + 3) + j] = -Etrx[3 * i + j];
      X[6 * (i + 3) + 3 + j] = Et[3 * i + j];
    }
}
static void crm(const real* v, const real* m, real* o) { /* motion cross: v x m */
  real a[3], b[3], c[3];
  cross3(v, m, a);
  cross3(v, m + 3, b);
  cross3(v + 3, m, c);
  for (int k = 0; k < 3; k++) { o[k] = a[k]; o[3 + k] = b[k] + c[k]; }
}
static void crf(const real* v, const real* f, real* o) { /* force cross: v x* f */
  real a[3], b[3], c[3];
  cross3(v, f, a);
  cross3(v + 3, f + 3, b);
  cross3(v, f + 3, c);
  for (int k = 0; k < 3; k++) { o[k] = a[k] + b[k]; o[3 + k] = c[k]; }
}
static void link_inertia6(const b2e_model* m, int i, real I[36]) {
  real ms = m->mass[i], c[3] = {m->com[i][0], m->com[i][1], m->com[i][2]};
  real cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0}, cxcx[9];
  m3_mul(cx, cx, cxcx);
  memset(I, 0, 36 * sizeof(real));
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      I[6 * a + b] = m->inertia[i][3 * a + b] - ms * cxcx[3 * a + b];
      I[6 * a + 3 + b] = ms * cx[3 * a + b];
      I[6 * (a + 3) + b] = -ms * cx[3 * a + b];
    }
  for (int a = 0; a < 3; a++) I[6 * (a + 3) + 3 + a] = ms;
}
static void mv6(const real* A, const real* x, real* y) {
  for (int i = 0; i < 6; i++) {
    real s = 0;
    for (int j = 0; j < 6; j++) s += A[6 * i + j] * x[j];
    y[i] = s;
  }
}
static real dot6(const real* a, const real* b) {
  real s = 0;
  for (int k = 0; k < 6; k++) s += a[k] * b[k];
  return s;
}

/* Articulated-body workspace kept after the forward-dynamics pass so that M^-1 * tau products
 * (Bullet: calcAccelerationDeltasMultiDof) cost one O(n) sweep each.                         */
typedef struct {
  real S[NL][6];
  real U[NL][6];
  real Dinv[NL];
  real IA[NL][36];
  real pA[NL][6];
  real v[NL][6];
  real c[NL][6];
  real u[NL];
} aba_t;

/* Featherstone ABA (RBDA Table 7.1) = Bullet computeAccelerationsArticulatedBodyAlgorithmMultiDof. */
static void aba(const b2e_model* m, const fk_t* fk, const real* qd, const real* tau, const real* grav, aba_t* w, real* qdd) {
  int n = m->n_links;
  for (int i = 0; i < n; i++) {
    int d = m->dof[i];
    real vp[6] = {0, 0, 0, 0, 0, 0};
    if (m->parent[i] >= 0) xform_motion(fk->E[i], fk->r[i], w->v[m->parent[i]], vp);
    for (int k = 0; k < 6; k++) w->S[i][k] = 0;
    if (d >= 0) {
      int off = m->jtype[i] == B2E_JOINT_REVOLUTE ? 0 : 3;
      for (int k = 0; k < 3; k++) w->S[i][off + k] = m->axis[i][k];
    }
    real vj[6];
    for (int k = 0; k < 6; k++) { vj[k] = d >= 0 ? w->S[i][k] * qd[d] : 0; w->v[i][k] = vp[k] + vj[k]; }
    crm(w->v[i], vj, w->c[i]);
    link_inertia6(m, i, w->IA[i]);
    real Iv[6];
    mv6(w->IA[i], w->v[i], Iv);
    crf(w->v[i], Iv, w->pA[i]);
  }
  for (int i = n - 1; i >= 0; i--) {
    int d = m->dof[i];
    real Ia[36], pa[6];
    memcpy(Ia, w->IA[i], sizeof(Ia));
    real Ic[6];
    mv6(w->IA[i], w->c[i], Ic);
    if (d >= 0) {
      mv6(w->IA[i], w->S[i], w->U[i]);
      real D = dot6(w->S[i], w->U[i]);
      w->Dinv[i] = 1 / D;
      w->u[i] = tau[d] - dot6(w->S[i], w->pA[i]);
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) Ia[6 * a + b] -= w->U[i][a] * w->Dinv[i] * w->U[i][b];
      real Iac[6];
      mv6(Ia, w->c[i], Iac);
      for (int k = 0; k < 6; k++) pa[k] = w->pA[i][k] + Iac[k] + w->U[i][k] * w->Dinv[i] * w->u[i];
    } else {
      for (int k = 0; k < 6; k++) pa[k] = w->pA[i][k] + Ic[k];
    }
    int p = m->parent[i];
    if (p >= 0) {
      real X[36], T[36];
      x6(fk->E[i], fk->r[i], X);
      /* IA_p += X^T Ia X */
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
          real s = 0;
          for (int k = 0; k < 6; k++) s += Ia[6 * a + k] * X[6 * k + b];
          T[6 * a + b] = s;
        }
      for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
          real s = 0;
          for (int k = 0; k < 6; k++) s += X[6 * k + a] * T[6 * k + b];
          w->IA[p][6 * a + b] += s;
        }
      real fp[6];
      xform_force_T(fk->E[i], fk->r[i], pa, fp);
      for (int k = 0; k < 6; k++) w->pA[p][k] += fp[k];
    }
  }
  real a[NL][6];
  real gb[3], g3[3] = {grav[0], grav[1], grav[2]}, Rb[9];
  for (int k = 0; k < 9; k++) Rb[k] = m->base_rot[k];
  m3t_vec(Rb, g3, gb);
  real a0[6] = {0, 0, 0, -gb[0], -gb[1], -gb[2]};
  for (int i = 0; i < n; i++) {
    int d = m->dof[i], p = m->parent[i];
    real ap[6];
    if (p >= 0) {
      xform_motion(fk->E[i], fk->r[i], a[p], ap);
    } else {
      /* link frame vs base frame: base is the parent with transform (E,r) */
      xform_motion(fk->E[i], fk->r[i], a0, ap);
    }
    for (int k = 0; k < 6; k++) ap[k] += w->c[i][k];
    if (d >= 0) {
      qdd[d] = w->Dinv[i] * (w->u[i] - dot6(w->U[i], ap));
      for (int k = 0; k < 6; k++) a[i][k] = ap[k] + w->S[i][k] * qdd[d];
    } else {
      for (int k = 0; k < 6; k++) a[i][k] = ap[k];
    }
  }
}

/* y = M^-1 * tau using the stored articulated inertias (Bullet calcAccelerationDeltasMultiDof). */
static void minv_mul(const b2e_model* m, const fk_t* fk, const aba_t* w, const real* tau, real* y) {
  int n = m->n_links;
  real p[NL][6], u[NL], a[NL][6];
  memset(p, 0, sizeof(p));
  for (int i = n - 1; i >= 0; i--) {
    int d = m->dof[i];
    real pa[6];
    if (d >= 0) {
      u[i] = tau[d] - dot6(w->S[i], p[i]);
      for (int k = 0; k < 6; k++) pa[k] = p[i][k] + w->U[i][k] * w->Dinv[i] * u[i];
    } else {
      u[i] = 0;
      for (int k = 0; k < 6; k++) pa[k] = p[i][k];
    }
    int pr = m->parent[i];
    if (pr >= 0) {
      real fp[6];
      xform_force_T(fk->E[i], fk->r[i], pa, fp);
      for (int k = 0; k < 6; k++) p[pr][k] += fp[k];
    }
  }
  for (int i = 0; i < n; i++) {
    int d = m->dof[i], pr = m->parent[i];
    real ap[6] = {0, 0, 0, 0, 0, 0};
    if (pr >= 0) xform_motion(fk->E[i], fk->r[i], a[pr], ap);
    if (d >= 0) {
      y[d] = w->Dinv[i] * (u[i] - dot6(w->U[i], ap));
      for (int k = 0; k < 6; k++) a[i][k] = ap[k] + w->S[i][k] * y[d];
    } else {
      for (int k = 0; k < 6; k++) a[i][k] = ap[k];
    }
  }
}

/* ------------------------------------------------------------------ collision
 * Broadphase: a static candidate-pair list (object x static world, robot proxies x object, robot proxies x static
 * world, robot self pairs) culled by bounding spheres.  Narrowphase: closed forms where they are exact (cube vertices on
 * the table top when the cube is wholly over it, sphere-box, sphere-plane, sphere-sphere), box-box SAT + face clipping
 * (include/b2env_narrowphase.h, shared with the CUDA kernel and pinned by tests/test_narrowphase.py) for everything
 * else: the cube at the table rim or against a leg, the finger-pad boxes against the cube and the table.            */
#define CT_CUBE_STATIC 0
#define CT_ARM_CUBE 1
#define CT_ARM_STATIC 2
#define CT_ARM_ARM 3

#define B2N_REAL real
#define B2N_FN static
#define B2N_SQRT(x) RSQRT(x)
#define B2N_FABS(x) RFABS(x)
#include "../include/b2env_narrowphase.h"

typedef struct {
  int key, type, link, link2;
  real pA[3], pB[3], n[3]; /* n points from B towards A; A = cube (cube-static) or the arm proxy (link); B = cube / world /
                              the second arm proxy (link2, self-collision) */
  real dist, mu, erp, cfm;
} contact_t;

typedef struct { contact_t* out; int nc, maxc, overflow; } clist_t;
static contact_t* push_contact(clist_t* L) {
  if (L->nc >= L->maxc) { L->overflow = 1; return NULL; }
  return &L->out[L->nc++];
}
static void sbox_get(const b2e_params* P, int k, real* c, real* h) {
  if (P->n_sboxes > 0) {
    for (int j = 0; j < 3; j++) { c[j] = P->sbox_c[k][j]; h[j] = P->sbox_h[k][j]; }
  } else {
    for (int j = 0; j < 3; j++) { c[j] = (real)0.5 * (P->table_min[j] + P->table_max[j]); h[j] = (real)0.5 * (P->table_max[j] - P->table_min[j]); }
  }
}
static const real IDENT3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

/* sphere (centre c, radius r) vs axis-aligned box: 1 if within margin; n from the box towards the sphere */
static int sphere_aabox(const real* c, real r, const real* bc, const real* bh, real margin, real* n, real* pB, real* dist, int* top) {
  real l[3], cl[3];
  int inside = 1;
  for (int j = 0; j < 3; j++) {
    l[j] = c[j] - bc[j];
    cl[j] = l[j] < -bh[j] ? -bh[j] : (l[j] > bh[j] ? bh[j] : l[j]);
    if (cl[j] != l[j]) inside = 0;
  }
  *top = 0;
  if (!inside) {
    real dv[3] = {l[0] - cl[0], l[1] - cl[1], l[2] - cl[2]};
    if (dv[0] == 0 && dv[1] == 0 && dv[2] > 0) { /* over the top face: the round-1 closed form, bit for bit */
      *dist = c[2] - r - (bc[2] + bh[2]);
      if (!(*dist < margin)) return 0;
      n[0] = 0; n[1] = 0; n[2] = 1;
      pB[0] = c[0]; pB[1] = c[1]; pB[2] = bc[2] + bh[2];
      *top = 1;
      return 1;
    }
    real d = RSQRT(dot3(dv, dv));
    *dist = d - r;
    if (!(*dist < margin)) return 0;
    for (int j = 0; j < 3; j++) { n[j] = dv[j] / d; pB[j] = bc[j] + cl[j]; }
    return 1;
  }
  int ax = 0;
  real best = bh[0] - RFABS(l[0]);
  for (int j = 1; j < 3; j++) {
    real pen = bh[j] - RFABS(l[j]);
    if (pen < best) { best = pen; ax = j; }
  }
  n[0] = n[1] = n[2] = 0;
  n[ax] = l[ax] >= 0 ? 1 : -1;
  cl[ax] = n[ax] * bh[ax];
  *dist = -best - r;
  for (int j = 0; j < 3; j++) pB[j] = bc[j] + cl[j];
  return 1;
}

static int collide(const b2e_model* m, const b2e_params* P, const fk_t* fk, const real* cpos, const real* cquat,
                   contact_t* out, int* overflow) {
  clist_t L = {out, 0, (P->max_contacts > 0 && P->max_contacts < MAXC) ? P->max_contacts : MAXC, 0};
  real Rc[9];
  quat_to_mat(cquat, Rc);
  const real a = P->cube_half, margin = P->contact_margin;
  const real ch[3] = {a, a, a};
  const real rb = a * (real)1.7320508075688772;   /* bounding sphere of the cube */
  const int nsb = P->n_sboxes > 0 ? P->n_sboxes : 1;
  real verts[8][3];
  for (int k = 0; k < 8; k++) {
    real l[3] = {(k & 1) ? a : -a, (k & 2) ? a : -a, (k & 4) ? a : -a};
    m3_vec(Rc, l, verts[k]);
    for (int j = 0; j < 3; j++) verts[k][j] += cpos[j];
  }
  /* ---- 1. cube vs static world ---- */
  const real top = P->table_max[2];
  const int cube_fast = cpos[0] - rb >= P->table_min[0] && cpos[0] + rb <= P->table_max[0] && cpos[1] - rb >= P->table_min[1] &&
                        cpos[1] + rb <= P->table_max[1] && cpos[2] >= top;
  if (cube_fast) { /* wholly over the table: vertex-face manifold (= box-box face clipping for a face-down cube) */
    for (int k = 0; k < 8; k++) {
      real dist = verts[k][2] - top;
      if (!(dist < margin)) continue;
      contact_t* c = push_contact(&L);
      if (!c) continue;
      c->key = B2E_KEY_CUBE_TABLE + k; c->type = CT_CUBE_STATIC; c->link = -1; c->link2 = -1;
      for (int j = 0; j < 3; j++) { c->pA[j] = verts[k][j]; c->pB[j] = verts[k][j]; }
      c->pB[2] = top;
      c->n[0] = 0; c->n[1] = 0; c->n[2] = 1;
      c->dist = dist; c->mu = P->cube_mu * P->table_mu; c->erp = P->erp; c->cfm = 0;
    }
  }
  for (int k = cube_fast ? 1 : 0; k < nsb; k++) { /* rim of the top slab, legs: general box-box behind a bounding-sphere cull */
    real bc[3], bh[3];
    sbox_get(P, k, bc, bh);
    real d[3] = {cpos[0] - bc[0], cpos[1] - bc[1], cpos[2] - bc[2]};
    if (RSQRT(dot3(d, d)) - (rb + RSQRT(dot3(bh, bh))) >= margin) continue;
    real n[3];
    b2n_contact pts[B2N_MAX_POINTS];
    int np = b2n_box_box(cpos, Rc, ch, bc, IDENT3, bh, margin, n, pts);
    for (int q = 0; q < np; q++) {
      contact_t* c = push_contact(&L);
      if (!c) continue;
      c->key = B2E_KEY_CUBE_SBOX + B2N_ID_STRIDE * k + pts[q].id; c->type = CT_CUBE_STATIC; c->link = -1; c->link2 = -1;
      for (int j = 0; j < 3; j++) { c->pA[j] = pts[q].pa[j]; c->pB[j] = pts[q].pb[j]; c->n[j] = n[j]; }
      c->dist = pts[q].dist; c->mu = P->cube_mu * (P->n_sboxes > 0 ? P->sbox_mu[k] : P->table_mu); c->erp = P->erp; c->cfm = 0;
    }
  }
  for (int k = 0; k < 8; k++) { /* ground plane z = 0 */
    if (cube_fast || !(verts[k][2] < margin)) continue;
    contact_t* c = push_contact(&L);
    if (!c) continue;
    c->key = B2E_KEY_CUBE_PLANE + k; c->type = CT_CUBE_STATIC; c->link = -1; c->link2 = -1;
    for (int j = 0; j < 3; j++) { c->pA[j] = verts[k][j]; c->pB[j] = verts[k][j]; }
    c->pB[2] = 0;
    c->n[0] = 0; c->n[1] = 0; c->n[2] = 1;
    c->dist = verts[k][2]; c->mu = P->cube_mu * P->plane_mu; c->erp = P->erp; c->cfm = 0;
  }
  /* ---- robot proxies in world coordinates ---- */
  real sc[B2E_MAX_SPHERES][3];
  for (int s = 0; s < m->n_spheres; s++) {
    int li = m->sph_link[s];
    real lc[3] = {m->sph_c[s][0], m->sph_c[s][1], m->sph_c[s][2]}, o[3];
    m3_vec(fk->R[li], lc, o);
    for (int j = 0; j < 3; j++) sc[s][j] = fk->p[li][j] + o[j];
  }
  real bxc[B2E_MAX_BOXES][3], bxh[B2E_MAX_BOXES][3], bxr[B2E_MAX_BOXES];
  for (int b = 0; b < m->n_boxes; b++) {
    int li = m->box_link[b];
    real lc[3] = {m->box_c[b][0], m->box_c[b][1], m->box_c[b][2]}, o[3];
    m3_vec(fk->R[li], lc, o);
    for (int j = 0; j < 3; j++) { bxc[b][j] = fk->p[li][j] + o[j]; bxh[b][j] = m->box_h[b][j]; }
    bxr[b] = RSQRT(dot3(bxh[b], bxh[b]));
  }
  b2n_shape caps[B2E_MAX_CAPS];
  real cap_e[B2E_MAX_CAPS][2][3];   /* world end points */
  for (int c = 0; c < m->n_caps; c++) {
    int li = m->cap_link[c];
    for (int e = 0; e < 2; e++) {
      real lc[3], o[3];
      for (int j = 0; j < 3; j++) lc[j] = e ? m->cap_p1[c][j] : m->cap_p0[c][j];
      m3_vec(fk->R[li], lc, o);
      for (int j = 0; j < 3; j++) cap_e[c][e][j] = fk->p[li][j] + o[j];
    }
    real ax[3] = {cap_e[c][1][0] - cap_e[c][0][0], cap_e[c][1][1] - cap_e[c][0][1], cap_e[c][1][2] - cap_e[c][0][2]};
    real len = RSQRT(dot3(ax, ax));
    caps[c].type = B2N_SEGMENT;
    for (int j = 0; j < 9; j++) caps[c].R[j] = 0;
    for (int j = 0; j < 3; j++) {
      caps[c].c[j] = (real)0.5 * (cap_e[c][0][j] + cap_e[c][1][j]);
      caps[c].R[3 * j + 2] = len > 0 ? ax[j] / len : (j == 2 ? (real)1 : (real)0);
      caps[c].h[j] = 0;
    }
    caps[c].h[2] = (real)0.5 * len;
    caps[c].r = m->cap_r[c];
  }
  /* ---- 2. robot vs cube: spheres (closest point on the box), finger-pad boxes (box-box), capsules (GJK / EPA) ---- */
  for (int s = 0; s < m->n_spheres; s++) {
    if (m->sph_flags[s] & 1) continue; /* self-collision partner only */
    real rel[3] = {sc[s][0] - cpos[0], sc[s][1] - cpos[1], sc[s][2] - cpos[2]}, l[3], cl[3];
    m3t_vec(Rc, rel, l);
    int inside = 1;
    for (int j = 0; j < 3; j++) {
      cl[j] = l[j] < -a ? -a : (l[j] > a ? a : l[j]);
      if (cl[j] != l[j]) inside = 0;
    }
    real nl[3], dist, r = m->sph_r[s];
    if (!inside) {
      real dv[3] = {l[0] - cl[0], l[1] - cl[1], l[2] - cl[2]};
      real d = RSQRT(dot3(dv, dv));
      dist = d - r;
      if (!(dist < margin)) continue;
      for (int j = 0; j < 3; j++) nl[j] = dv[j] / d;
    } else {
      /* centre inside the box: push out through the nearest face */
      int ax = 0;
      real best = a - RFABS(l[0]);
      for (int j = 1; j < 3; j++) {
        real pen = a - RFABS(l[j]);
        if (pen < best) { best = pen; ax = j; }
      }
      nl[0] = nl[1] = nl[2] = 0;
      nl[ax] = l[ax] >= 0 ? 1 : -1;
      cl[ax] = nl[ax] * a;
      dist = -best - r;
    }
    contact_t* c = push_contact(&L);
    if (!c) continue;
    c->key = B2E_KEY_SPHERE_CUBE + s; c->type = CT_ARM_CUBE; c->link = m->sph_link[s]; c->link2 = -1;
    real nw[3], pw[3];
    m3_vec(Rc, nl, nw);
    m3_vec(Rc, cl, pw);
    for (int j = 0; j < 3; j++) {
      c->n[j] = nw[j];
      c->pB[j] = cpos[j] + pw[j];
      c->pA[j] = sc[s][j] - nw[j] * r;
    }
    c->dist = dist; c->mu = P->cube_mu * m->sph_mu[s];
    c->erp = m->sph_erp[s] >= 0 ? m->sph_erp[s] : P->erp;
    c->cfm = m->sph_cfm[s];
  }
  for (int b = 0; b < m->n_boxes; b++) {
    real d[3] = {bxc[b][0] - cpos[0], bxc[b][1] - cpos[1], bxc[b][2] - cpos[2]};
    if (RSQRT(dot3(d, d)) - (rb + bxr[b]) >= margin) continue;
    real n[3];
    b2n_contact pts[B2N_MAX_POINTS];
    int np = b2n_box_box(bxc[b], fk->R[m->box_link[b]], bxh[b], cpos, Rc, ch, margin, n, pts);
    for (int q = 0; q < np; q++) {
      contact_t* c = push_contact(&L);
      if (!c) continue;
      c->key = B2E_KEY_BOX_CUBE + B2N_ID_STRIDE * b + pts[q].id; c->type = CT_ARM_CUBE; c->link = m->box_link[b]; c->link2 = -1;
      for (int j = 0; j < 3; j++) { c->pA[j] = pts[q].pa[j]; c->pB[j] = pts[q].pb[j]; c->n[j] = n[j]; }
      c->dist = pts[q].dist; c->mu = P->cube_mu * m->box_mu[b];
      c->erp = m->box_erp[b] >= 0 ? m->box_erp[b] : P->erp;
      c->cfm = m->box_cfm[b];
    }
  }
  for (int c = 0; c < m->n_caps; c++) {
    real d[3] = {caps[c].c[0] - cpos[0], caps[c].c[1] - cpos[1], caps[c].c[2] - cpos[2]};
    if (RSQRT(dot3(d, d)) - (rb + caps[c].h[2] + caps[c].r) >= margin) continue;
    b2n_shape cube;
    cube.type = B2N_BOX; cube.r = 0;
    for (int j = 0; j < 3; j++) { cube.c[j] = cpos[j]; cube.h[j] = a; }
    for (int j = 0; j < 9; j++) cube.R[j] = Rc[j];
    real n[3];
    b2n_contact pt;
    if (!b2n_convex_contact(&caps[c], &cube, margin, n, &pt)) continue;
    contact_t* ct = push_contact(&L);
    if (!ct) continue;
    ct->key = B2E_KEY_CAP_CUBE + c; ct->type = CT_ARM_CUBE; ct->link = m->cap_link[c]; ct->link2 = -1;
    for (int j = 0; j < 3; j++) { ct->pA[j] = pt.pa[j]; ct->pB[j] = pt.pb[j]; ct->n[j] = n[j]; }
    ct->dist = pt.dist; ct->mu = P->cube_mu * m->cap_mu[c]; ct->erp = P->erp; ct->cfm = 0;
  }
  /* ---- 3. robot vs static world ---- */
  for (int s = 0; s < m->n_spheres; s++) {
    if (m->sph_flags[s] & 1) continue;
    real r = m->sph_r[s];
    int kept = 0;
    for (int k = 0; k < nsb && kept < 3; k++) { /* at most three static boxes per sphere */
      real bc[3], bh[3], n[3], pB[3], dist;
      int is_top;
      sbox_get(P, k, bc, bh);
      real d[3] = {sc[s][0] - bc[0], sc[s][1] - bc[1], sc[s][2] - bc[2]};
      if (RSQRT(dot3(d, d)) - (r + RSQRT(dot3(bh, bh))) >= margin) continue;
      if (!sphere_aabox(sc[s], r, bc, bh, margin, n, pB, &dist, &is_top)) continue;
      kept++;
      contact_t* c = push_contact(&L);
      if (!c) continue;
      c->key = (k == 0 && is_top) ? B2E_KEY_SPHERE_TABLE + s : B2E_KEY_SPHERE_SBOX + 8 * s + k;
      c->type = CT_ARM_STATIC; c->link = m->sph_link[s]; c->link2 = -1;
      for (int j = 0; j < 3; j++) { c->n[j] = n[j]; c->pB[j] = pB[j]; c->pA[j] = sc[s][j] - n[j] * r; }
      c->dist = dist; c->mu = (P->n_sboxes > 0 ? P->sbox_mu[k] : P->table_mu) * m->sph_mu[s];
      c->erp = m->sph_erp[s] >= 0 ? m->sph_erp[s] : P->erp;
      c->cfm = m->sph_cfm[s];
    }
  }
  for (int s = 0; s < m->n_spheres; s++) { /* ground plane */
    real r = m->sph_r[s], dist = sc[s][2] - r;
    if ((m->sph_flags[s] & 1) || !(dist < margin)) continue;
    contact_t* c = push_contact(&L);
    if (!c) continue;
    c->key = B2E_KEY_SPHERE_PLANE + s; c->type = CT_ARM_STATIC; c->link = m->sph_link[s]; c->link2 = -1;
    c->n[0] = 0; c->n[1] = 0; c->n[2] = 1;
    for (int j = 0; j < 3; j++) { c->pA[j] = sc[s][j]; c->pB[j] = sc[s][j]; }
    c->pA[2] = dist; c->pB[2] = 0;
    c->dist = dist; c->mu = P->plane_mu * m->sph_mu[s];
    c->erp = m->sph_erp[s] >= 0 ? m->sph_erp[s] : P->erp;
    c->cfm = m->sph_cfm[s];
  }
  /* finger-pad boxes: family by family (the CUDA kernel compacts each family with one ballot): vertices on the table top
   * (pad wholly over the table), general box-box against the static boxes, vertices on the ground plane */
  int bfast[B2E_MAX_BOXES];
  real bv[B2E_MAX_BOXES][8][3];
  for (int b = 0; b < m->n_boxes; b++) {
    const real* Rb = fk->R[m->box_link[b]];
    bfast[b] = bxc[b][0] - bxr[b] >= P->table_min[0] && bxc[b][0] + bxr[b] <= P->table_max[0] &&
               bxc[b][1] - bxr[b] >= P->table_min[1] && bxc[b][1] + bxr[b] <= P->table_max[1] && bxc[b][2] >= top;
    for (int k = 0; k < 8; k++) {
      real l[3] = {(k & 1) ? bxh[b][0] : -bxh[b][0], (k & 2) ? bxh[b][1] : -bxh[b][1], (k & 4) ? bxh[b][2] : -bxh[b][2]};
      m3_vec(Rb, l, bv[b][k]);
      for (int j = 0; j < 3; j++) bv[b][k][j] += bxc[b][j];
    }
  }
  for (int b = 0; b < m->n_boxes; b++) {
    const real erp = m->box_erp[b] >= 0 ? m->box_erp[b] : P->erp;
    for (int k = 0; k < 8; k++) {
      real dist = bv[b][k][2] - top;
      if (!bfast[b] || !(dist < margin)) continue;
      contact_t* c = push_contact(&L);
      if (!c) continue;
      c->key = B2E_KEY_BOXV_TABLE + 8 * b + k; c->type = CT_ARM_STATIC; c->link = m->box_link[b]; c->link2 = -1;
      for (int j = 0; j < 3; j++) { c->pA[j] = bv[b][k][j]; c->pB[j] = bv[b][k][j]; }
      c->pB[2] = top;
      c->n[0] = 0; c->n[1] = 0; c->n[2] = 1;
      c->dist = dist; c->mu = P->table_mu * m->box_mu[b]; c->erp = erp; c->cfm = m->box_cfm[b];
    }
  }
  for (int b = 0; b < m->n_boxes; b++) {
    const real* Rb = fk->R[m->box_link[b]];
    const real erp = m->box_erp[b] >= 0 ? m->box_erp[b] : P->erp;
    for (int k = bfast[b] ? 1 : 0; k < nsb; k++) {
      real bc[3], bh[3];
      sbox_get(P, k, bc, bh);
      real d[3] = {bxc[b][0] - bc[0], bxc[b][1] - bc[1], bxc[b][2] - bc[2]};
      if (RSQRT(dot3(d, d)) - (bxr[b] + RSQRT(dot3(bh, bh))) >= margin) continue;
      real n[3];
      b2n_contact pts[B2N_MAX_POINTS];
      int np = b2n_box_box(bxc[b], Rb, bxh[b], bc, IDENT3, bh, margin, n, pts);
      for (int q = 0; q < np; q++) {
        contact_t* c = push_contact(&L);
        if (!c) continue;
        c->key = B2E_KEY_BOX_SBOX + B2N_ID_STRIDE * (8 * b + k) + pts[q].id; c->type = CT_ARM_STATIC; c->link = m->box_link[b]; c->link2 = -1;
        for (int j = 0; j < 3; j++) { c->pA[j] = pts[q].pa[j]; c->pB[j] = pts[q].pb[j]; c->n[j] = n[j]; }
        c->dist = pts[q].dist; c->mu = (P->n_sboxes > 0 ? P->sbox_mu[k] : P->table_mu) * m->box_mu[b]; c->erp = erp; c->cfm = m->box_cfm[b];
      }
    }
  }
  for (int b = 0; b < m->n_boxes; b++) {
    const real erp = m->box_erp[b] >= 0 ? m->box_erp[b] : P->erp;
    for (int k = 0; k < 8; k++) { /* ground plane */
      if (bfast[b] || !(bv[b][k][2] < margin)) continue;
      contact_t* c = push_contact(&L);
      if (!c) continue;
      c->key = B2E_KEY_BOXV_PLANE + 8 * b + k; c->type = CT_ARM_STATIC; c->link = m->box_link[b]; c->link2 = -1;
      for (int j = 0; j < 3; j++) { c->pA[j] = bv[b][k][j]; c->pB[j] = bv[b][k][j]; }
      c->pB[2] = 0;
      c->n[0] = 0; c->n[1] = 0; c->n[2] = 1;
      c->dist = bv[b][k][2]; c->mu = P->plane_mu * m->box_mu[b]; c->erp = erp; c->cfm = m->box_cfm[b];
    }
  }
  /* capsules vs static boxes (GJK / EPA behind an axis-aligned bounding-box cull), vs the ground plane (end spheres) */
  for (int c = 0; c < m->n_caps; c++) {
    for (int k = 0; k < nsb; k++) {
      real bc[3], bh[3];
      sbox_get(P, k, bc, bh);
      int apart = 0;
      for (int j = 0; j < 3; j++) {
        real lo = cap_e[c][0][j] < cap_e[c][1][j] ? cap_e[c][0][j] : cap_e[c][1][j];
        real hi = cap_e[c][0][j] < cap_e[c][1][j] ? cap_e[c][1][j] : cap_e[c][0][j];
        if (lo - caps[c].r - (bc[j] + bh[j]) >= margin || (bc[j] - bh[j]) - (hi + caps[c].r) >= margin) apart = 1;
      }
      if (apart) continue;
      b2n_shape box;
      box.type = B2N_BOX; box.r = 0;
      for (int j = 0; j < 3; j++) { box.c[j] = bc[j]; box.h[j] = bh[j]; }
      for (int j = 0; j < 9; j++) box.R[j] = IDENT3[j];
      real n[3];
      b2n_contact pt;
      if (!b2n_convex_contact(&caps[c], &box, margin, n, &pt)) continue;
      contact_t* ct = push_contact(&L);
      if (!ct) continue;
      ct->key = B2E_KEY_CAP_SBOX + 8 * c + k; ct->type = CT_ARM_STATIC; ct->link = m->cap_link[c]; ct->link2 = -1;
      for (int j = 0; j < 3; j++) { ct->pA[j] = pt.pa[j]; ct->pB[j] = pt.pb[j]; ct->n[j] = n[j]; }
      ct->dist = pt.dist; ct->mu = (P->n_sboxes > 0 ? P->sbox_mu[k] : P->table_mu) * m->cap_mu[c]; ct->erp = P->erp; ct->cfm = 0;
    }
  }
  for (int c = 0; c < m->n_caps; c++) {
    for (int e = 0; e < 2; e++) {
      real dist = cap_e[c][e][2] - caps[c].r;
      if (!(dist < margin)) continue;
      contact_t* ct = push_contact(&L);
      if (!ct) continue;
      ct->key = B2E_KEY_CAP_PLANE + 2 * c + e; ct->type = CT_ARM_STATIC; ct->link = m->cap_link[c]; ct->link2 = -1;
      ct->n[0] = 0; ct->n[1] = 0; ct->n[2] = 1;
      for (int j = 0; j < 3; j++) { ct->pA[j] = cap_e[c][e][j]; ct->pB[j] = cap_e[c][e][j]; }
      ct->pA[2] = dist; ct->pB[2] = 0;
      ct->dist = dist; ct->mu = P->plane_mu * m->cap_mu[c]; ct->erp = P->erp; ct->cfm = 0;
    }
  }
  /* ---- 4. robot self-collision (URDF_USE_SELF_COLLISION, panda_env.py:53): sphere pairs of non-neighbouring links ---- */
  for (int p = 0; p < m->n_self_pairs; p++) {
    const int sa = m->self_a[p], sb = m->self_b[p];
    real d[3] = {sc[sa][0] - sc[sb][0], sc[sa][1] - sc[sb][1], sc[sa][2] - sc[sb][2]};
    real dn = RSQRT(dot3(d, d));
    real dist = dn - (m->sph_r[sa] + m->sph_r[sb]);
    if (!(dist < margin) || !(dn > (real)1e-9)) continue;
    contact_t* c = push_contact(&L);
    if (!c) continue;
    c->key = B2E_KEY_SELF + p; c->type = CT_ARM_ARM; c->link = m->sph_link[sa]; c->link2 = m->sph_link[sb];
    for (int j = 0; j < 3; j++) {
      c->n[j] = d[j] / dn;
      c->pA[j] = sc[sa][j] - c->n[j] * m->sph_r[sa];
      c->pB[j] = sc[sb][j] + c->n[j] * m->sph_r[sb];
    }
    c->dist = dist; c->mu = m->sph_mu[sa] * m->sph_mu[sb]; c->erp = P->erp; c->cfm = 0;
  }
  *overflow = L.overflow;
  return L.nc;
}

/* btPlaneSpace1 [EXT-recalled]: deterministic tangent basis of a unit normal */
static void plane_space(const real* n, real* p, real* q) {
  if (RFABS(n[2]) > (real)0.7071067811865475244) {
    real a = n[1] * n[1] + n[2] * n[2];
    real k = 1 / RSQRT(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    real a = n[0] * n[0] + n[1] * n[1];
    real k = 1 / RSQRT(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

/* ------------------------------------------------------------------ constraint rows */
#define ROW_MOTOR 0
#define ROW_LIMIT 1
#define ROW_NORMAL 2
#define ROW_FRICTION 3
#define ISL_ARM 0
#define ISL_CUBE 1

typedef struct {
  int type, island, nrow; /* nrow: index of the normal row (friction rows) */
  real J[NV], W[NV];
  real rhs, cfm, diag, lo, hi, lam, mu;
} row_t;

typedef struct {
  const b2e_model* m;
  const b2e_params* P;
  const fk_t* fk;
  const aba_t* w;
  int nd;
  real cinv_m, cinv_I;
} rowctx_t;

static void finish_row(const rowctx_t* cx, row_t* r, const real* vstar, real desired) {
  int nd = cx->nd;
  minv_mul(cx->m, cx->fk, cx->w, r->J, r->W);
  for (int k = 0; k < 3; k++) {
    r->W[nd + k] = r->J[nd + k] * cx->cinv_m;
    r->W[nd + 3 + k] = r->J[nd + 3 + k] * cx->cinv_I;
  }
  real d = 0, jv = 0;
  for (int k = 0; k < nd + 6; k++) { d += r->J[k] * r->W[k]; jv += r->J[k] * vstar[k]; }
  r->diag = d + r->cfm;
  r->rhs = desired - jv;
  r->lam = 0;
}

/* contact direction row: relative velocity of A w.r.t. B along dir */
static void contact_J(const rowctx_t* cx, const contact_t* c, const real* cpos, const real* dir, real* J) {
  int nd = cx->nd;
  for (int k = 0; k < nd + 6; k++) J[k] = 0;
  if (c->type == CT_CUBE_STATIC) {
    real rel[3] = {c->pA[0] - cpos[0], c->pA[1] - cpos[1], c->pA[2] - cpos[2]}, t[3];
    cross3(rel, dir, t);
    for (int k = 0; k < 3; k++) { J[nd + k] = dir[k]; J[nd + 3 + k] = t[k]; }
  } else {
    real Jl[ND][3], Ja[ND][3];
    point_jacobian(cx->m, cx->fk, c->link, c->pA, Jl, Ja);
    for (int d = 0; d < nd; d++) J[d] = dot3(Jl[d], dir);
    if (c->type == CT_ARM_ARM) { /* self-collision: velocity of the point on link minus that of the point on link2 */
      point_jacobian(cx->m, cx->fk, c->link2, c->pB, Jl, Ja);
      for (int d = 0; d < nd; d++) J[d] -= dot3(Jl[d], dir);
    }
    if (c->type == CT_ARM_CUBE) {
      real rel[3] = {c->pB[0] - cpos[0], c->pB[1] - cpos[1], c->pB[2] - cpos[2]}, t[3];
      cross3(rel, dir, t);
      for (int k = 0; k < 3; k++) { J[nd + k] = -dir[k]; J[nd + 3 + k] = -t[k]; }
    }
  }
}

/* ------------------------------------------------------------------ observation, reward */
static void ee_state(const b2e_model* m, const fk_t* fk, const real* qd, real* pos, real* quat, real* vlin) {
  int ee = m->ee_link;
  real c[3] = {m->com[ee][0], m->com[ee][1], m->com[ee][2]}, o[3];
  m3_vec(fk->R[ee], c, o);
  for (int k = 0; k < 3; k++) pos[k] = fk->p[ee][k] + o[k];
  mat_to_quat(fk->R[ee], quat);
  real Jl[ND][3], Ja[ND][3];
  point_jacobian(m, fk, ee, pos, Jl, Ja);
  for (int k = 0; k < 3; k++) {
    real s = 0;
    for (int d = 0; d < m->n_dof; d++) s += Jl[d][k] * qd[d];
    vlin[k] = s;
  }
}

/* panda_push_gym_env.py:150-187 (Push, 33 entries) / panda_reach_gym_env.py:140-171 (Reach, 30) */
static int extended_observation(const b2e_model* m, const b2e_params* P, const fk_t* fk, const real* q, const real* qd,
                                const real* cpos, const real* cquat, const real* target, real* obs) {
  real pos[3], quat[4], vl[3], eu[3];
  ee_state(m, fk, qd, pos, quat, vl);
  quat_to_euler(quat, eu);
  int n = 0;
  for (int k = 0; k < 3; k++) obs[n++] = pos[k];
  for (int k = 0; k < 3; k++) obs[n++] = eu[k];
  for (int k = 0; k < 3; k++) obs[n++] = (vl[k] - P->vel_mean[k]) / P->vel_std[k]; /* panda_env.py:171-181 */
  if (P->n_obs_joints > 0) { /* icub_env.py:242-247: the controlled joints only */
    for (int k = 0; k < P->n_obs_joints; k++) obs[n++] = q[P->obs_dof[k]];
  } else {
    for (int d = 0; d < m->n_dof; d++) obs[n++] = q[d];
  }
  real ceu[3];
  quat_to_euler(cquat, ceu);
  for (int k = 0; k < 3; k++) obs[n++] = cpos[k];
  for (int k = 0; k < 3; k++) obs[n++] = ceu[k];
  /* object pose in the hand frame; both orientations round-trip through Euler (:168-174) */
  real hq[4], oq[4], hqi[4], rel[3], relq[4], releu[3], d[3];
  euler_to_quat(eu, hq);
  euler_to_quat(ceu, oq);
  hqi[0] = -hq[0]; hqi[1] = -hq[1]; hqi[2] = -hq[2]; hqi[3] = hq[3];
  for (int k = 0; k < 3; k++) d[k] = cpos[k] - pos[k];
  quat_rot(hqi, d, rel);
  quat_mul(hqi, oq, relq);
  quat_to_euler(relq, releu);
  for (int k = 0; k < 3; k++) obs[n++] = rel[k];
  for (int k = 0; k < 3; k++) obs[n++] = releu[k];
  if (P->task == B2E_TASK_PUSH || P->task == B2E_TASK_GRASP)
    for (int k = 0; k < 3; k++) obs[n++] = target[k];
  return n;
}

static real dist3(const real* a, const real* b) {
  real d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
  return RSQRT(dot3(d, d));
}


/* ------------------------------------------------------------------ damped-least-squares IK
 * p.calculateInverseKinematics(robot, ee, pos, orn, maxNumIterations=100, residualThreshold=1e-3)
 * (panda_env.py:269-272) restated [EXT-recalled: IKTrajectoryHelper, IK2_VEL_DLS_WITH_ORIENTATION]:
 * start from the current joint positions; while |p_t - p| > residual and it < max: FK, 6 x n Jacobian of
 * the EE link frame, e = [p_t - p ; rotation vector of q_t * q^-1], dq = J^T (J J^T + lambda I)^-1 e,
 * largest joint step clamped to 45 deg, q += dq.  Returns positions of every movable joint.        */
static void quat_conj_mul(const real* a, const real* b, real* o) { /* o = a * conj(b) */
  real c[4] = {-b[0], -b[1], -b[2], b[3]};
  quat_mul(a, c, o);
}
static int solve6(real U[6][7]) { /* Gaussian elimination with partial pivoting on [U | e] */
  for (int k = 0; k < 6; k++) {
    int piv = k;
    for (int r = k + 1; r < 6; r++) if (RFABS(U[r][k]) > RFABS(U[piv][k])) piv = r;
    if (piv != k) for (int c = 0; c < 7; c++) { real t = U[k][c]; U[k][c] = U[piv][c]; U[piv][c] = t; }
    real inv = 1 / U[k][k];
    for (int r = k + 1; r < 6; r++) {
      real f = U[r][k] * inv;
      for (int c = k; c < 7; c++) U[r][c] -= f * U[k][c];
    }
  }
  for (int k = 5; k >= 0; k--) {
    real s = U[k][6];
    for (int c = k + 1; c < 6; c++) s -= U[k][c] * U[c][6];
    U[k][6] = s / U[k][k];
  }
  return 0;
}
static int ik_dls(const b2e_model* m, const b2e_params* P, const real* q0, const real* tpos, const real* tquat, real* qout) {
  int nd = m->n_dof, ee = m->ee_link, it;
  real q[ND];
  for (int d = 0; d < nd; d++) q[d] = q0[d];
  for (it = 0; it < P->ik_iters; it++) {
    fk_t fk;
    forward_kinematics(m, q, &fk);
    real dp[3] = {tpos[0] - fk.p[ee][0], tpos[1] - fk.p[ee][1], tpos[2] - fk.p[ee][2]};
    if (RSQRT(dot3(dp, dp)) <= P->ik_residual) break;
    real cq[4], eq[4], er[3];
    mat_to_quat(fk.R[ee], cq);
    quat_conj_mul(tquat, cq, eq);
    if (eq[3] < 0) { eq[0] = -eq[0]; eq[1] = -eq[1]; eq[2] = -eq[2]; eq[3] = -eq[3]; }
    real vn = RSQRT(eq[0] * eq[0] + eq[1] * eq[1] + eq[2] * eq[2]);
    if (vn > (real)1e-9) {
      real ang = 2 * RATAN2(vn, eq[3]);
      for (int k = 0; k < 3; k++) er[k] = eq[k] / vn * ang;
    } else er[0] = er[1] = er[2] = 0;
    real Jl[ND][3], Ja[ND][3];
    point_jacobian(m, &fk, ee, fk.p[ee], Jl, Ja);
    real e[6] = {dp[0], dp[1], dp[2], er[0], er[1], er[2]};
    real U[6][7];
    for (int r = 0; r < 6; r++) {
      for (int c = 0; c < 6; c++) {
        real sacc = 0;
        for (int d = 0; d < nd; d++) {
          real jr = r < 3 ? Jl[d][r] : Ja[d][r - 3], jc = c < 3 ? Jl[d][c] : Ja[d][c - 3];
          sacc += jr * jc;
        }
        U[r][c] = sacc + (r == c ? P->ik_damping : 0);
      }
      U[r][6] = e[r];
    }
    solve6(U);
    real dq[ND], mx = 0;
    for (int d = 0; d < nd; d++) {
      real sacc = 0;
      for (int r = 0; r < 6; r++) sacc += (r < 3 ? Jl[d][r] : Ja[d][r - 3]) * U[r][6];
      dq[d] = sacc;
      if (RFABS(sacc) > mx) mx = RFABS(sacc);
    }
    real scale = mx > (real)0.78539816339 ? (real)0.78539816339 / mx : 1;
    for (int d = 0; d < nd; d++) q[d] += dq[d] * scale;
  }
  for (int d = 0; d < nd; d++) qout[d] = q[d];
  return it;
}

/* ------------------------------------------------------------------ one physics step of one env */
typedef struct {
  real q[ND], qd[ND], cpos[3], cquat[4], cv[3], cw[3], mtarget[ND], hand_pose[6];
  int cache_key[B2E_CACHE_SLOTS];
  real cache_lam[B2E_CACHE_SLOTS][3];
  int flags, iters, n_contacts, n_rows;
  real contact_out[MAXC][8];
} env_t;

static void physics_step(const b2e_model* m, const b2e_params* P, env_t* e, int kp_ctrl_active, int grip_active, int ik_active) {
  int nd = m->n_dof, nv = nd + 6;
  real dt = P->dt;
  fk_t fk;
  aba_t w;
  forward_kinematics(m, e->q, &fk);

  /* --- unconstrained velocities: v* = v + dt * a(q, v) --- */
  real tau[ND], qdd[ND], grav[3] = {P->gravity[0], P->gravity[1], P->gravity[2]};
  for (int d = 0; d < nd; d++) tau[d] = -m->joint_damping[d] * e->qd[d];
  aba(m, &fk, e->qd, tau, grav, &w, qdd);
  real vstar[NV];
  for (int d = 0; d < nd; d++) vstar[d] = e->qd[d] + dt * qdd[d];
  real vn = RSQRT(dot3(e->cv, e->cv)), wn = RSQRT(dot3(e->cw, e->cw));
  for (int k = 0; k < 3; k++) {
    vstar[nd + k] = e->cv[k] + dt * (grav[k] - e->cv[k] * (P->damp_lin_k1 + P->damp_lin_k2 * vn));
    vstar[nd + 3 + k] = e->cw[k] + dt * (-e->cw[k] * (P->damp_ang_k1 + P->damp_ang_k2 * wn));
  }

  /* --- collision detection on the pre-step poses --- */
  contact_t con[MAXC];
  int overflow = 0;
  int nc = collide(m, P, &fk, e->cpos, e->cquat, con, &overflow);
  if (overflow) e->flags |= B2E_ST_CONTACT_OVERFLOW;

  /* --- rows: motors, limits, contact normals, contact frictions --- */
  row_t rows[MAXROWS];
  int nr = 0;
  rowctx_t cx = {m, P, &fk, &w, nd, 1 / P->cube_mass, 1 / P->cube_inertia};
  for (int d = 0; d < nd; d++) { /* btMultiBodyJointMotor: desired = kp*(target-q)/dt (kd=1, erp=1) */
    row_t* r = &rows[nr++];
    memset(r, 0, sizeof(*r));
    r->type = ROW_MOTOR; r->island = ISL_ARM;
    r->J[d] = 1;
    real kp = (kp_ctrl_active && is_ctrl_dof(P, d)) ? P->kp_ctrl : ((grip_active && d >= P->n_ctrl) ? P->kp_grip : P->kp_hold);
    real mv = m->max_vel[d];
    if (ik_active && P->ik_max_vel > 0 && (is_ctrl_dof(P, d) || P->n_obs_joints > 0)) { kp = P->kp_ik_max_vel; mv = P->ik_max_vel; } /* panda_env.py:285-291 */
    real desired = kp * (e->mtarget[d] - e->q[d]) / dt;
    if (mv > 0) {
      if (desired > mv) desired = mv;
      if (desired < -mv) desired = -mv;
    }
    r->cfm = 0;
    r->hi = m->max_force[d] * dt;
    r->lo = -r->hi;
    finish_row(&cx, r, vstar, desired);
  }
  int nlim = 0;
  for (int d = 0; d < nd; d++) { /* btMultiBodyJointLimitConstraint rows, only near a limit */
    for (int side = 0; side < 2; side++) {
      real dist = side == 0 ? e->q[d] - m->lower[d] : m->upper[d] - e->q[d];
      if (!(dist < m->limit_margin[d])) continue;
      if (nlim >= B2E_MAX_LIMROWS) { e->flags |= B2E_ST_LIMIT_OVERFLOW; continue; }
      nlim++;
      row_t* r = &rows[nr++];
      memset(r, 0, sizeof(*r));
      r->type = ROW_LIMIT; r->island = ISL_ARM;
      r->J[d] = side == 0 ? 1 : -1;
      real pen = dist + P->slop;
      real desired = pen > 0 ? -pen / dt : -pen * P->erp / dt;
      r->lo = 0; r->hi = (real)1e30; r->cfm = 0;
      finish_row(&cx, r, vstar, desired);
    }
  }
  int coupled = 0;
  int normal_row[MAXC];
  for (int c = 0; c < nc; c++) {
    row_t* r = &rows[nr];
    memset(r, 0, sizeof(*r));
    r->type = ROW_NORMAL;
    r->island = con[c].type == CT_CUBE_STATIC ? ISL_CUBE : ISL_ARM;
    if (con[c].type == CT_ARM_CUBE) coupled = 1;
    contact_J(&cx, &con[c], e->cpos, con[c].n, r->J);
    real pen = con[c].dist + P->slop;
    real desired = pen > 0 ? -pen / dt : -pen * con[c].erp / dt;
    r->lo = 0; r->hi = (real)1e30; r->cfm = con[c].cfm;
    finish_row(&cx, r, vstar, desired);
    normal_row[c] = nr++;
  }
  for (int c = 0; c < nc; c++) {
    real t1[3], t2[3];
    plane_space(con[c].n, t1, t2);
    for (int f = 0; f < 2; f++) {
      row_t* r = &rows[nr++];
      memset(r, 0, sizeof(*r));
      r->type = ROW_FRICTION;
      r->island = rows[normal_row[c]].island;
      r->nrow = normal_row[c];
      r->mu = con[c].mu;
      contact_J(&cx, &con[c], e->cpos, f == 0 ? t1 : t2, r->J);
      r->cfm = 0;
      finish_row(&cx, r, vstar, 0);
    }
  }

  /* --- warm start from the cache (contact rows only) --- */
  real dv[NV];
  for (int k = 0; k < nv; k++) dv[k] = 0;
  for (int c = 0; c < nc; c++) {
    for (int s = 0; s < B2E_CACHE_SLOTS; s++) {
      if (e->cache_key[s] != con[c].key) continue;
      int ri[3] = {normal_row[c], nd + nlim + nc + 2 * c, nd + nlim + nc + 2 * c + 1};
      for (int j = 0; j < 3; j++) {
        real l = e->cache_lam[s][j] * P->warmstart;
        rows[ri[j]].lam = l;
        for (int k = 0; k < nv; k++) dv[k] += rows[ri[j]].W[k] * l;
      }
      break;
    }
  }

  /* --- projected Gauss-Seidel (Bullet btMultiBodyConstraintSolver::solveSingleIteration order:
   *     non-contact rows, contact normals, then frictions bounded by mu * normal impulse);
   *     islands that do not share a body converge independently --- */
  int done_isl[2] = {0, 0};
  int it = 0;
  int has_cube_rows = 0;
  for (int i = 0; i < nr; i++) if (rows[i].island == ISL_CUBE) has_cube_rows = 1;
  if (!has_cube_rows) done_isl[ISL_CUBE] = 1;
  for (it = 0; it < P->solver_iters; it++) {
    real res[2] = {0, 0};
    for (int i = 0; i < nr; i++) {
      row_t* r = &rows[i];
      int isl = coupled ? 0 : r->island;
      if (done_isl[isl]) continue;
      if (r->type == ROW_FRICTION) {
        real lim = r->mu * rows[r->nrow].lam;
        r->lo = -lim; r->hi = lim;
      }
      real jdv = 0;
      for (int k = 0; k < nv; k++) jdv += r->J[k] * dv[k];
      real dl = (r->rhs - jdv - r->cfm * r->lam) / r->diag;
      real nl = r->lam + dl;
      if (nl < r->lo) nl = r->lo;
      if (nl > r->hi) nl = r->hi;
      dl = nl - r->lam;
      r->lam = nl;
      for (int k = 0; k < nv; k++) dv[k] += r->W[k] * dl;
      real rv = dl * r->diag;
      if (rv * rv > res[isl]) res[isl] = rv * rv;
    }
    if (coupled) {
      if (res[0] <= P->residual_tol) { it++; break; }
    } else {
      if (!done_isl[0] && res[0] <= P->residual_tol) done_isl[0] = 1;
      if (!done_isl[1] && res[1] <= P->residual_tol) done_isl[1] = 1;
      if (done_isl[0] && done_isl[1]) { it++; break; }
    }
  }
  e->iters = it;
  e->n_contacts = nc;
  e->n_rows = nr;

  /* --- integrate (semi-implicit Euler) --- */
  for (int d = 0; d < nd; d++) {
    e->qd[d] = vstar[d] + dv[d];
    e->q[d] += dt * e->qd[d];
  }
  for (int k = 0; k < 3; k++) {
    e->cv[k] = vstar[nd + k] + dv[nd + k];
    e->cw[k] = vstar[nd + 3 + k] + dv[nd + 3 + k];
    e->cpos[k] += dt * e->cv[k];
  }
  {
    real wl = RSQRT(dot3(e->cw, e->cw)), ang = wl * dt, f;
    if (ang < (real)1e-3) f = (real)0.5 * dt - dt * dt * dt * (real)(1.0 / 48.0) * wl * wl;
    else f = RSIN((real)0.5 * ang) / wl;
    real dq[4] = {e->cw[0] * f, e->cw[1] * f, e->cw[2] * f, RCOS((real)0.5 * ang)}, nq[4];
    quat_mul(dq, e->cquat, nq);
    real nn = 1 / RSQRT(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
    for (int k = 0; k < 4; k++) e->cquat[k] = nq[k] * nn;
  }

  /* --- write back the contact cache + diagnostics --- */
  for (int s = 0; s < B2E_CACHE_SLOTS; s++) { e->cache_key[s] = -1; e->cache_lam[s][0] = e->cache_lam[s][1] = e->cache_lam[s][2] = 0; }
  memset(e->contact_out, 0, sizeof(e->contact_out));
  for (int c = 0; c < nc; c++) {
    int ri[3] = {normal_row[c], nd + nlim + nc + 2 * c, nd + nlim + nc + 2 * c + 1};
    e->cache_key[c] = con[c].key;
    for (int j = 0; j < 3; j++) e->cache_lam[c][j] = rows[ri[j]].lam;
    e->contact_out[c][0] = (real)con[c].key;
    e->contact_out[c][1] = con[c].dist;
    for (int j = 0; j < 3; j++) { e->contact_out[c][2 + j] = con[c].n[j]; e->contact_out[c][5 + j] = rows[ri[j]].lam; }
  }
  int bad = 0;
  for (int d = 0; d < nd; d++) if (!isfinite(e->q[d]) || !isfinite(e->qd[d])) bad = 1;
  for (int k = 0; k < 3; k++) if (!isfinite(e->cpos[k]) || !isfinite(e->cv[k]) || !isfinite(e->cw[k])) bad = 1;
  if (bad) e->flags |= B2E_ST_NAN;
}

/* ------------------------------------------------------------------ batched entry points */
static void load_env(const b2e_model* m, const b2o_state* S, int b, env_t* e) {
  int nd = m->n_dof;
  for (int d = 0; d < nd; d++) {
    e->q[d] = S->q[b * nd + d];
    e->qd[d] = S->qd[b * nd + d];
    e->mtarget[d] = S->mtarget[b * nd + d];
  }
  for (int k = 0; k < 3; k++) {
    e->cpos[k] = S->obj_pose[b * 7 + k];
    e->cv[k] = S->obj_vel[b * 6 + k];
    e->cw[k] = S->obj_vel[b * 6 + 3 + k];
  }
  for (int k = 0; k < 4; k++) e->cquat[k] = S->obj_pose[b * 7 + 3 + k];
  for (int s = 0; s < B2E_CACHE_SLOTS; s++) {
    e->cache_key[s] = S->cache_key[b * B2E_CACHE_SLOTS + s];
    for (int j = 0; j < 3; j++) e->cache_lam[s][j] = S->cache_lam[(b * B2E_CACHE_SLOTS + s) * 3 + j];
  }
  for (int k = 0; k < 6; k++) e->hand_pose[k] = S->hand_pose[b * 6 + k];
  e->flags = S->status[b * 4];
  e->iters = 0; e->n_contacts = 0; e->n_rows = 0;
}
static void store_env(const b2e_model* m, b2o_state* S, int b, const env_t* e) {
  int nd = m->n_dof;
  for (int d = 0; d < nd; d++) {
    S->q[b * nd + d] = (float)e->q[d];
    S->qd[b * nd + d] = (float)e->qd[d];
    S->mtarget[b * nd + d] = (float)e->mtarget[d];
  }
  for (int k = 0; k < 3; k++) {
    S->obj_pose[b * 7 + k] = (float)e->cpos[k];
    S->obj_vel[b * 6 + k] = (float)e->cv[k];
    S->obj_vel[b * 6 + 3 + k] = (float)e->cw[k];
  }
  for (int k = 0; k < 4; k++) S->obj_pose[b * 7 + 3 + k] = (float)e->cquat[k];
  for (int s = 0; s < B2E_CACHE_SLOTS; s++) {
    S->cache_key[b * B2E_CACHE_SLOTS + s] = e->cache_key[s];
    for (int j = 0; j < 3; j++) S->cache_lam[(b * B2E_CACHE_SLOTS + s) * 3 + j] = (float)e->cache_lam[s][j];
  }
  for (int k = 0; k < 6; k++) S->hand_pose[b * 6 + k] = (float)e->hand_pose[k];
  S->status[b * 4 + 0] = e->flags;
  S->status[b * 4 + 1] = e->iters;
  S->status[b * 4 + 2] = e->n_contacts;
  S->status[b * 4 + 3] = e->n_rows;
  for (int c = 0; c < MAXC; c++)
    for (int j = 0; j < 8; j++) S->contacts[(b * MAXC + c) * 8 + j] = (float)e->contact_out[c][j];
}

static void step_env(const b2e_model* m, const b2e_params* P, b2o_state* S, int b, const float* action, float* obs,
                     float* reward, float* done, int nsub, int mode) {
  env_t e;
  load_env(m, S, b, &e);
  int counter = S->counters[b * 2], terminated = S->counters[b * 2 + 1];
  real target[3] = {S->target[b * 3], S->target[b * 3 + 1], S->target[b * 3 + 2]};
  real act[ND];
  if (mode == B2E_MODE_ACTION)
    for (int k = 0; k < P->n_act; k++) act[k] = action[b * P->n_act + k];
  for (int sub = 0; sub < nsub; sub++) {
    if (mode == B2E_MODE_ACTION && !P->use_ik) {
      /* panda_push_gym_env.py:225-230: action *= 0.05 (compounding in place across repeats, quirk E.4),
       * new = q[:n_ctrl] + action; panda_env.py:303: clamp to [ll, ul] */
      for (int k = 0; k < P->n_ctrl; k++) { /* iCub: icub_push_gym_env.py:256-257, icub_env.py:347-361 */
        const int d = ctrl_dof_of(P, k);
        act[k] *= P->act_scale;
        real t = e.q[d] + act[k];
        if (t < m->lower[d]) t = m->lower[d];
        if (t > m->upper[d]) t = m->upper[d];
        e.mtarget[d] = t;
      }
      if (P->task == B2E_TASK_GRASP) { /* gripper command in [-1,1] -> both finger targets in [0, 0.04] */
        real g = action[b * P->n_act + P->n_ctrl];
        for (int k = P->n_ctrl; k < m->n_dof; k++) {
          real t = (real)0.02 + (real)0.02 * g;
          if (t < m->lower[k]) t = m->lower[k];
          if (t > m->upper[k]) t = m->upper[k];
          e.mtarget[k] = t;
        }
      }
    }
    if ((mode == B2E_MODE_ACTION || mode == B2E_MODE_IK_POSE) && P->use_ik) {
      /* task level (panda_push_gym_env.py:197-222): hand_pose += scaled action, clamp to rotation limits
       * and workspace; robot level (panda_env.py:229-282): z re-clamp, orientation = home euler when it is
       * not controlled, IK, position targets for every movable joint with gain 0.2 */
      if (mode == B2E_MODE_ACTION) {
        if (!P->ik_orientation) {
          for (int k = 0; k < 3; k++) { act[k] *= P->act_scale_pos; e.hand_pose[k] += act[k]; }
        } else {
          for (int k = 0; k < 3; k++) { act[k] *= P->act_scale_pos; e.hand_pose[k] += act[k]; }
          for (int k = 3; k < 6; k++) {
            act[k] *= P->act_scale_rot;
            real v = e.hand_pose[k] + act[k];
            if (v < P->eu_lim[k - 3][0]) v = P->eu_lim[k - 3][0];
            if (v > P->eu_lim[k - 3][1]) v = P->eu_lim[k - 3][1];
            e.hand_pose[k] = v;
          }
        }
        for (int k = 0; k < 3; k++) {
          if (e.hand_pose[k] < P->ws_lim[k][0]) e.hand_pose[k] = P->ws_lim[k][0];
          if (e.hand_pose[k] > P->ws_lim[k][1]) e.hand_pose[k] = P->ws_lim[k][1];
        }
      }
      real tp[3] = {e.hand_pose[0], e.hand_pose[1], e.hand_pose[2]}, eu[3], tq[4];
      if (tp[2] < P->ws_lim[2][0]) tp[2] = P->ws_lim[2][0];
      if (tp[2] > P->ws_lim[2][1]) tp[2] = P->ws_lim[2][1];
      for (int k = 0; k < 3; k++) { /* robot level: clamp to the rotation limits (panda_env.py:250-256,
                                       icub_env.py:284-289), or the home orientation when it is not controlled */
        eu[k] = P->ik_orientation ? e.hand_pose[3 + k] : P->home_hand_pose[3 + k];
        if (P->ik_orientation) {
          if (eu[k] < P->eu_lim[k][0]) eu[k] = P->eu_lim[k][0];
          if (eu[k] > P->eu_lim[k][1]) eu[k] = P->eu_lim[k][1];
        }
      }
      euler_to_quat(eu, tq);
      { /* COM pose -> link-frame pose of the hand (icub_env.py:303-305; zero offset for the Panda) */
        real off[3] = {P->ik_link_offset[0], P->ik_link_offset[1], P->ik_link_offset[2]}, o[3];
        quat_rot(tq, off, o);
        for (int k = 0; k < 3; k++) tp[k] += o[k];
      }
      real qik[ND];
      ik_dls(m, P, e.q, tp, tq, qik);
      for (int d = 0; d < m->n_dof; d++) /* blocked joints keep the rest pose (icub_env.py:314-317) */
        e.mtarget[d] = (P->n_obs_joints > 0 && !is_ctrl_dof(P, d)) ? (real)m->home[d] : qik[d];
    }
    physics_step(m, P, &e, mode != B2E_MODE_HOLD && mode != B2E_MODE_IK_POSE && !P->use_ik, mode != B2E_MODE_HOLD && P->task == B2E_TASK_GRASP,
                 mode == B2E_MODE_IK_POSE || (mode == B2E_MODE_ACTION && P->use_ik));
    if (mode == B2E_MODE_ACTION) {
      /* _termination() inside apply_action (:239-242): counter only advances when not terminated */
      fk_t fk;
      forward_kinematics(m, e.q, &fk);
      real d;
      if (P->task == B2E_TASK_PUSH) d = dist3(e.cpos, target);
      else if (P->task == B2E_TASK_GRASP) d = (e.cpos[2] - P->grasp_rest_z >= P->grasp_lift) ? 0 : 2 * P->dist_min + 1;
      else { real pos[3], qt[4], vl[3]; ee_state(m, &fk, e.qd, pos, qt, vl); d = dist3(pos, e.cpos); }
      int term = 0;
      if (P->goal_env) { if (counter > P->max_steps) term = 1; }   /* panda_push_gym_goal_env.py:106-110 */
      else if (d <= P->dist_min) { terminated = 1; term = 1; }
      else if (terminated || counter > P->max_steps) term = 1;
      if (term) break;
      counter++;
    }
  }
  store_env(m, S, b, &e);
  S->counters[b * 2] = counter;
  if (mode == B2E_MODE_ACTION || obs) {
    fk_t fk;
    forward_kinematics(m, e.q, &fk);
    real raw[B2E_MAX_OBS];
    int n = extended_observation(m, P, &fk, e.q, e.qd, e.cpos, e.cquat, target, raw);
    for (int k = 0; k < n; k++) {
      S->raw_obs[b * P->n_obs + k] = (float)raw[k];
      if (obs) /* scale_gym_data, utils.py:91 (no clipping) */
        obs[b * P->n_obs + k] = (float)(2 * ((raw[k] - P->obs_low[k]) / (P->obs_high[k] - P->obs_low[k])) - 1);
    }
    /* _termination (:301-316) then _compute_reward (:318-331) */
    real d1 = dist3(raw, e.cpos), d2 = 0, rew;
    int dn = 0;
    if (P->task == B2E_TASK_PUSH && P->goal_env) { /* GoalEnv variants of both robots (panda_push_gym_goal_env.py:96-122,
                                                     icub_push_gym_goal_env.py:99-132): done = _termination() or
                                                     is_success; reward = -(d > dist_min) */
      d2 = dist3(e.cpos, target);
      dn = (counter > P->max_steps) || (d2 <= P->dist_min);
      rew = d2 > P->dist_min ? -1 : 0;
    } else if (P->reward_kind == B2E_REWARD_ICUB_REACH) { /* icub_reach_gym_env.py:301-330: the bonus is ADDED */
      if (d1 <= P->dist_min) { terminated = 1; dn = 1; }
      else if (terminated || counter > P->max_steps) dn = 1;
      rew = -d1;
      if (d1 <= P->dist_min) rew += (real)1000.0 + (100 - d1 * 80);
    } else if (P->reward_kind == B2E_REWARD_ICUB_PUSH0 || P->reward_kind == B2E_REWARD_ICUB_PUSH1) {
      /* icub_push_gym_env.py:327-373 */
      d2 = dist3(e.cpos, target);
      if (d2 <= P->dist_min) { terminated = 1; dn = 1; }
      else if (terminated || counter > P->max_steps) dn = 1;
      if (P->reward_kind == B2E_REWARD_ICUB_PUSH0) rew = -d1 - d2;
      else {
        real d0 = S->shaping[b * 2], dmax = S->shaping[b * 2 + 1];
        rew = (real)0.125 * (1 - d1 / d0);
        if (!(d1 > (real)0.1)) rew += (real)0.25 * (1 - d2 / dmax);
      }
      if (d2 <= P->dist_min) rew += (real)1000.0;
    } else if (P->task == B2E_TASK_PUSH) {
      d2 = dist3(e.cpos, target);
      if (P->goal_env) { /* done = _termination() or is_success; reward = -(d > dist_min) (:96-122) */
        dn = (counter > P->max_steps) || (d2 <= P->dist_min);
        rew = d2 > P->dist_min ? -1 : 0;
      } else {
        if (d2 <= P->dist_min) { terminated = 1; dn = 1; }
        else if (terminated || counter > P->max_steps) dn = 1;
        rew = -d1 - d2;
        if (d2 <= P->dist_min) rew = (real)1000.0 + (100 - d2 * 80);
      }
    } else if (P->task == B2E_TASK_GRASP) {
      /* PandaGrasp (new task): reach the object, close the fingers, lift it grasp_lift above its rest height */
      real lift = e.cpos[2] - P->grasp_rest_z;
      int success = lift >= P->grasp_lift;
      if (success) { terminated = 1; dn = 1; }
      else if (terminated || counter > P->max_steps) dn = 1;
      real l = lift < 0 ? 0 : (lift > P->grasp_lift ? P->grasp_lift : lift);
      rew = -d1 + 50 * l;
      if (success) rew = (real)1000.0;
    } else {
      if (d1 <= P->dist_min) { terminated = 1; dn = 1; }
      else if (terminated || counter > P->max_steps) dn = 1;
      rew = -d1;
      if (d1 <= P->dist_min) rew = (real)1000.0 + (100 - d1 * 80);
    }
    if (reward) reward[b] = (float)rew;
    if (done) done[b] = (float)dn;
  }
  if (mode != B2E_MODE_OBSERVE) S->counters[b * 2 + 1] = terminated;
}

#include <pthread.h>
typedef struct {
  const b2e_model* m; const b2e_params* P; b2o_state* S; const float* action; float* obs; float* reward; float* done;
  int nsub, mode, b0, b1;
} job_t;
static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (int b = j->b0; b < j->b1; b++) step_env(j->m, j->P, j->S, b, j->action, j->obs, j->reward, j->done, j->nsub, j->mode);
  return NULL;
}

/* One env.step() for every env; nthreads > 1 splits the batch over host threads. */
int b2o_step(const b2e_model* m, const b2e_params* P, b2o_state* S, const float* action, float* obs, float* reward,
             float* done, int nsub, int mode, int nthreads) {
  if (!m || !P || !S) return B2E_EINVAL;
  if (m->n_dof + 6 > NV) return B2E_EINVAL;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads == 1 || S->B < 2 * nthreads) {
    job_t j = {m, P, S, action, obs, reward, done, nsub, mode, 0, S->B};
    worker(&j);
    return 0;
  }
  pthread_t th[256];
  job_t jobs[256];
  int per = (S->B + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    int b0 = t * per, b1 = b0 + per > S->B ? S->B : b0 + per;
    if (b0 > b1) b0 = b1;
    jobs[t] = (job_t){m, P, S, action, obs, reward, done, nsub, mode, b0, b1};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return 0;
}

/* Masked reset (panda_env.py:63-79 home state; world_env.py:81-84 object pose). */
int b2o_reset(const b2e_model* m, const b2e_params* P, b2o_state* S, const uint8_t* mask, const float* obj_init_pose,
              const float* target) {
  (void)P;
  int nd = m->n_dof;
  for (int b = 0; b < S->B; b++) {
    if (mask && !mask[b]) continue;
    for (int d = 0; d < nd; d++) {
      S->q[b * nd + d] = m->home[d];
      S->qd[b * nd + d] = 0;
      S->mtarget[b * nd + d] = m->home[d];
    }
    for (int k = 0; k < 7; k++) S->obj_pose[b * 7 + k] = obj_init_pose[b * 7 + k];
    for (int k = 0; k < 6; k++) S->obj_vel[b * 6 + k] = 0;
    for (int k = 0; k < 3; k++) S->target[b * 3 + k] = target[b * 3 + k];
    S->counters[b * 2] = 0; S->counters[b * 2 + 1] = 0;
    for (int s = 0; s < B2E_CACHE_SLOTS; s++) {
      S->cache_key[b * B2E_CACHE_SLOTS + s] = -1;
      for (int j = 0; j < 3; j++) S->cache_lam[(b * B2E_CACHE_SLOTS + s) * 3 + j] = 0;
    }
    for (int k = 0; k < 6; k++) S->hand_pose[b * 6 + k] = P->home_hand_pose[k];
    for (int k = 0; k < 4; k++) S->status[b * 4 + k] = 0;
  }
  return 0;
}

/* ------------------------------------------------------------------ known-answer helpers (tests only) */
int b2o_fk(const b2e_model* m, const float* q, float* link_pos /*[n][3]*/, float* link_rot /*[n][9]*/) {
  real qq[ND] = {0};
  for (int d = 0; d < m->n_dof; d++) qq[d] = q[d];
  fk_t fk;
  forward_kinematics(m, qq, &fk);
  for (int i = 0; i < m->n_links; i++) {
    for (int k = 0; k < 3; k++) link_pos[3 * i + k] = (float)fk.p[i][k];
    for (int k = 0; k < 9; k++) link_rot[9 * i + k] = (float)fk.R[i][k];
  }
  return 0;
}
int b2o_forward_dynamics(const b2e_model* m, const b2e_params* P, const float* q, const float* qd, const float* tau,
                         float* qdd) {
  real qq[ND] = {0}, qv[ND] = {0}, tt[ND] = {0}, out[ND], g[3] = {P->gravity[0], P->gravity[1], P->gravity[2]};
  for (int d = 0; d < m->n_dof; d++) { qq[d] = q[d]; qv[d] = qd[d]; tt[d] = tau[d]; }
  fk_t fk; aba_t w;
  forward_kinematics(m, qq, &fk);
  aba(m, &fk, qv, tt, g, &w, out);
  for (int d = 0; d < m->n_dof; d++) qdd[d] = (float)out[d];
  return 0;
}
int b2o_minv(const b2e_model* m, const float* q, float* Minv /*[nd][nd]*/) {
  real qq[ND] = {0}, z[ND] = {0}, out[ND], g[3] = {0, 0, 0};
  int nd = m->n_dof;
  for (int d = 0; d < nd; d++) { qq[d] = q[d]; z[d] = 0; }
  fk_t fk; aba_t w;
  forward_kinematics(m, qq, &fk);
  aba(m, &fk, z, z, g, &w, out);
  for (int c = 0; c < nd; c++) {
    real e[ND];
    for (int d = 0; d < nd; d++) e[d] = d == c;
    minv_mul(m, &fk, &w, e, out);
    for (int d = 0; d < nd; d++) Minv[d * nd + c] = (float)out[d];
  }
  return 0;
}
int b2o_ee_jacobian(const b2e_model* m, const float* q, float* J /*[6][nd] linear rows then angular rows*/,
                    float* pos, float* quat) {
  real qq[ND] = {0}, z[ND] = {0};
  int nd = m->n_dof;
  for (int d = 0; d < nd; d++) { qq[d] = q[d]; z[d] = 0; }
  fk_t fk;
  forward_kinematics(m, qq, &fk);
  real p[3], qt[4], vl[3];
  ee_state(m, &fk, z, p, qt, vl);
  real Jl[ND][3], Ja[ND][3];
  point_jacobian(m, &fk, m->ee_link, p, Jl, Ja);
  for (int d = 0; d < nd; d++)
    for (int k = 0; k < 3; k++) { J[k * nd + d] = (float)Jl[d][k]; J[(3 + k) * nd + d] = (float)Ja[d][k]; }
  for (int k = 0; k < 3; k++) pos[k] = (float)p[k];
  for (int k = 0; k < 4; k++) quat[k] = (float)qt[k];
  return 0;
}
int b2o_ik(const b2e_model* m, const b2e_params* P, const float* q0, const float* tpos, const float* tquat, float* qout) {
  real q[ND] = {0}, tp[3], tq[4], o[ND];
  for (int d = 0; d < m->n_dof; d++) q[d] = q0[d];
  for (int k = 0; k < 3; k++) tp[k] = tpos[k];
  for (int k = 0; k < 4; k++) tq[k] = tquat[k];
  int it = ik_dls(m, P, q, tp, tq, o);
  for (int d = 0; d < m->n_dof; d++) qout[d] = (float)o[d];
  return it;
}
/* box-box narrowphase (include/b2env_narrowphase.h) on its own: out = [n][8] = pa(3) pb(3) dist id, normal[3] */
int b2o_box_box(const float* cA, const float* RA, const float* hA, const float* cB, const float* RB, const float* hB, float margin,
                float* normal, float* out) {
  real a[3], ra[9], ha[3], b[3], rb[9], hb[3], n[3];
  for (int k = 0; k < 3; k++) { a[k] = cA[k]; ha[k] = hA[k]; b[k] = cB[k]; hb[k] = hB[k]; }
  for (int k = 0; k < 9; k++) { ra[k] = RA[k]; rb[k] = RB[k]; }
  b2n_contact pts[B2N_MAX_POINTS];
  int np = b2n_box_box(a, ra, ha, b, rb, hb, (real)margin, n, pts);
  for (int q = 0; q < np; q++) {
    for (int k = 0; k < 3; k++) { out[8 * q + k] = (float)pts[q].pa[k]; out[8 * q + 3 + k] = (float)pts[q].pb[k]; }
    out[8 * q + 6] = (float)pts[q].dist;
    out[8 * q + 7] = (float)pts[q].id;
  }
  if (np > 0) for (int k = 0; k < 3; k++) normal[k] = (float)n[k];
  return np;
}
/* GJK / EPA contact of two rounded convex shapes (include/b2env_narrowphase.h) on its own.  shape = [type, c(3), R(9), h(3), r]
 * as 17 floats; out = pa(3) pb(3) dist id, normal[3]; returns 1 when within margin                                        */
int b2o_convex_contact(const float* shapeA, const float* shapeB, float margin, float* normal, float* out) {
  b2n_shape S[2];
  const float* src[2] = {shapeA, shapeB};
  for (int i = 0; i < 2; i++) {
    S[i].type = (int)src[i][0];
    for (int k = 0; k < 3; k++) { S[i].c[k] = src[i][1 + k]; S[i].h[k] = src[i][13 + k]; }
    for (int k = 0; k < 9; k++) S[i].R[k] = src[i][4 + k];
    S[i].r = src[i][16];
  }
  real n[3];
  b2n_contact c;
  if (!b2n_convex_contact(&S[0], &S[1], (real)margin, n, &c)) return 0;
  for (int k = 0; k < 3; k++) { out[k] = (float)c.pa[k]; out[3 + k] = (float)c.pb[k]; normal[k] = (float)n[k]; }
  out[6] = (float)c.dist;
  out[7] = (float)c.id;
  return 1;
}
/* the contact set of one configuration: out = [n][16] = key type link link2 pA(3) pB(3) n(3) dist mu erp */
int b2o_collide(const b2e_model* m, const b2e_params* P, const float* q, const float* obj_pose, float* out, int* overflow) {
  real qq[ND] = {0}, cp[3], cq[4];
  for (int d = 0; d < m->n_dof; d++) qq[d] = q[d];
  for (int k = 0; k < 3; k++) cp[k] = obj_pose[k];
  for (int k = 0; k < 4; k++) cq[k] = obj_pose[3 + k];
  fk_t fk;
  forward_kinematics(m, qq, &fk);
  contact_t con[MAXC];
  int nc = collide(m, P, &fk, cp, cq, con, overflow);
  for (int c = 0; c < nc; c++) {
    float* o = out + 16 * c;
    o[0] = (float)con[c].key; o[1] = (float)con[c].type; o[2] = (float)con[c].link; o[3] = (float)con[c].link2;
    for (int k = 0; k < 3; k++) { o[4 + k] = (float)con[c].pA[k]; o[7 + k] = (float)con[c].pB[k]; o[10 + k] = (float)con[c].n[k]; }
    o[13] = (float)con[c].dist; o[14] = (float)con[c].mu; o[15] = (float)con[c].erp;
  }
  return nc;
}
int b2o_real_size(void) { return (int)sizeof(real); }
