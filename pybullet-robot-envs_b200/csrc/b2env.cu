// b2env.cu — B200 (sm_100a) batched rigid-body step + reward pipeline and its C-ABI.
//
// Sixteen lanes advance one environment through a complete env.step() of the reference
// (pybullet_robot_envs/envs/panda_envs/panda_push_gym_env.py:244-255): action -> motor targets
// (:225-230, panda_env.py:303-310), p.stepSimulation (:236) = articulated forward dynamics +
// collision detection + PGS over motor/limit/contact rows + semi-implicit Euler, then the
// observation gather (:150-187), termination (:301-316) and reward (:318-331), all fused in a
// single launch.  See DESIGN.md for the algorithm statement shared with oracle/b2oracle.c.
//
// Mapping of the work onto the machine: ONE ENVIRONMENT = 16 LANES, two environments per warp.  Every
// stage needs <= 16 lanes (12 links, 9 dofs, 8 cube vertices, 14 spheres, 15 velocity components, 16
// rows per set); all warp collectives carry the member mask of the environment's own half-warp, so the
// two environments of a warp may diverge and otherwise share every instruction.
//   * kinematics / dynamics : lane = link; parent->child recursions are carried by pointer-jumping
//     over __shfl_sync (log2(depth) rounds), child->parent accumulations by a host-built shuffle
//     schedule (heavy-path suffix scans); the joint-space inertia matrix is built CRBA-style in
//     world coordinates and inverted in registers by Gauss-Jordan (lane = row).
//   * collision             : lane = cube vertex / lane = collision sphere, ballot compaction.
//   * PGS                   : lane = constraint row.  The solver runs on the Delassus form
//     A = J M^-1 J^T, so one row update is a handful of scalar instructions + one shuffle + one
//     LDS/FMA per lane instead of two n-wide dot/axpy.  Motor rows are never materialised (their
//     Delassus block is M^-1 itself) and, when they form an island of their own, a Gauss-Seidel sweep
//     over them is evaluated as the affine map lambda' = G lambda + c.  This is the same Gauss-Seidel
//     iteration Bullet runs in delta-velocity space (identical iterates in exact arithmetic, same
//     row order, same exit test).
// There is no dense contraction anywhere (largest matrix 9x9 / 48x48 per env): no tensor cores.
//
// State lives in HBM as env-major field groups ([B][width] arrays, see enum b2e_field), read
// once and written once per step.

#ifndef B2E_EMU   // tools/emu compiles this file with g++ and cuda_emu.h force-included (CPU debugging of kernel logic)
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b2env.h"

#ifdef B2E_EMU
__thread unsigned char smem_raw[232448] __attribute__((aligned(128)));
#define B2E_LAUNCH(kern, grid, block, smem, stream, ...) emu::launch(dim3(grid), dim3(block), (smem), [=]() { kern(__VA_ARGS__); })
#else
#define B2E_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

#define FULL 0xffffffffu
#ifdef PROFILE_WARM   // probe build: per code site, visits of a warp and visits with fewer than 32 active lanes (b2e_debug_lane_sites)
__device__ unsigned long long g_lane_site[16][2];
#define LANE_SITE(k) do { const unsigned am_ = __activemask(); if ((int)(threadIdx.x & 31) == __ffs(am_) - 1) { \
    atomicAdd(&g_lane_site[k][0], 1ull); if (am_ != FULL) atomicAdd(&g_lane_site[k][1], 1ull); } } while (0)
#else
#define LANE_SITE(k) do { } while (0)
#endif
#ifndef PHASE_SYNC
#define PHASE_SYNC 1       // block barriers between stages keep the warps of a block on the same code
#endif
#if PHASE_SYNC
#define PHASE_BARRIER() __syncthreads()
#else
#define PHASE_BARRIER() __syncwarp()
#endif
#ifndef WPB
#define WPB 8             // warps per block (two envs per warp)
#endif
#ifndef MINB
#define MINB 2            // resident blocks per SM the register allocation targets (16 warps = 32 envs, 128 regs)
#endif
#define TLMAX 12          // links handled by the group kernel (transforms staged in shared memory; < 15: lane 15 is the zero lane)
#define GMAX 48           // generic rows per env (3 limit + 36 contact rows), in sets of 16
#define BIGS 40           // row stride of a big constraint system (<= 39 generic rows) in an overflow slot
#ifndef NSLOT
#define NSLOT 2           // overflow slots per block for environments with more than 16 generic rows
#endif
#define WSTRIDE 16        // row stride of the W = M^-1 J^T table
// The solver's sweep loop must stay SMALL: a sweep-capped environment walks it 150 times, alone or next to other heavy warps,
// and once the loop bodies of the warps on an SM outgrow the instruction caches (L0 ~6 KB per scheduler, L1.5 32 KB per SM)
// every row update waits for instruction fetches (measured with four heavy warps per tail block: 550-640 cycles per row
// update with the row loops unrolled by 2 or 4, 120-130 with rolled loops).  Row loops are therefore NOT unrolled.
#ifndef SWEEP_UNROLL
#define SWEEP_UNROLL 1
#endif
#define B2E_PRAGMA(x) _Pragma(#x)
#define B2E_UNROLL(n) B2E_PRAGMA(unroll n)
// Cost-ordered scheduling (full-batch launches of the group kernel).  The environments of a block advance in phase
// (block barriers between the stages), so a block lasts as long as its slowest environment, and a launch as long as its
// slowest block: with contiguous blocks, one jammed contact (all solver sweeps over ~45 rows, ~0.4 ms) held 15 ordinary
// environments and half an SM's registers / shared memory for the whole launch.  Instead, every environment files itself,
// at the end of a step, under a cost class estimated from its own solve (sweeps x rows); the NEXT full-batch step is
//   * a TAIL launch on the library's own high-priority stream: the heavy class (big constraint systems, or an estimated
//     cost above SCHED_TAIL_COST), in lean blocks (TAIL_WPB warps, one overflow slot per environment) that start first, and
//   * the MAIN launch: everything else, slot -> environment taken from the class lists, heaviest class first, so that
//     blocks are cost-homogeneous.
// Both run concurrently; results do not depend on which slot steps an environment (tests/test_emu_kernels.py).
#define NBK_MAIN 16                         // cost classes of the main launch
#define NBK (NBK_MAIN + 1)                  // + the tail class
#define TAIL_CAP 1024                       // environments the tail launch can take (overflow: heaviest main class)
#ifndef SCHED_TAIL_COST
#define SCHED_TAIL_COST 160000              // estimated solve cycles that send an environment to the tail launch
#endif
#define SCHED_CNT(q, b) ((q) * NBK + (b))                                        // int sched[]: 4 x NBK counters (ring),
#define SCHED_LIST(par, b, B) (4 * NBK + ((par) * NBK_MAIN + (b)) * (B))         // 2 x NBK_MAIN class lists of capacity B,
#define SCHED_TAIL(par, B) (4 * NBK + 2 * NBK_MAIN * (B) + (par) * TAIL_CAP)      // 2 x tail list
#define SCHED_INTS(B) (4 * NBK + 2 * NBK_MAIN * (B) + 2 * TAIL_CAP)
#define ROLE_PLAIN 0                        // slot -> environment: identity / env_ids (subset launches, no scheduling)
#define ROLE_MAIN 1
#define ROLE_TAIL 2
#define SCRATCH_PER_ENV (BIGS * BIGS + BIGS * WSTRIDE + NDMAX * BIGS)   // A | W | W^T(arm part), same layout as a slot
#define SCRATCH_PAD 64
#define NDMAX 9           // dofs handled by the warp kernel (Panda: 7 arm + 2 fingers)
#define NLMAX 32          // links (lanes)

#define CT_CUBE_STATIC 0     // object vs static world: rows act on the object only
#define CT_ARM_CUBE 1        // robot proxy (point on `link`) vs object
#define CT_ARM_STATIC 2      // robot proxy vs static world
#define CT_ARM_ARM 3         // robot self-collision: point on `link` vs point on `link2`
#define ROW_MOTOR 0
#define ROW_LIMIT 1
#define ROW_NORMAL 2
#define ROW_FRICTION 3

// ------------------------------------------------------------------------------------------
// device-side model (preprocessed copy of b2e_model)
struct DevModel {
  int n_links, n_dof, ee_link, n_spheres, fk_rounds;
  int parent[NLMAX], jtype[NLMAX], dof[NLMAX];
  int dof_link[NLMAX];
  unsigned link_dofmask[NLMAX];  // dofs on the path base -> link (inclusive)
  unsigned acc_sched[2][NLMAX];  // child->parent accumulation schedule: 4-bit source lane per round (15 = none)
  signed char tsched[24][NLMAX]; // tree kernel (b2env_tree.cuh): source lane per round, -1 = none
  float jpos[NLMAX][3], jrot[NLMAX][9], axis[NLMAX][3];
  float mass[NLMAX], com[NLMAX][3], inertia[NLMAX][9];
  float lower[NLMAX], upper[NLMAX], limit_margin[NLMAX], max_force[NLMAX], max_vel[NLMAX], joint_damping[NLMAX],
      home[NLMAX];
  float base_pos[3], base_rot[9];
  int sph_link[B2E_MAX_SPHERES];
  float sph_c[B2E_MAX_SPHERES][3], sph_r[B2E_MAX_SPHERES], sph_mu[B2E_MAX_SPHERES], sph_erp[B2E_MAX_SPHERES],
      sph_cfm[B2E_MAX_SPHERES];
  int n_boxes, n_self_pairs;
  int box_link[B2E_MAX_BOXES];
  float box_c[B2E_MAX_BOXES][3], box_h[B2E_MAX_BOXES][3], box_mu[B2E_MAX_BOXES], box_erp[B2E_MAX_BOXES], box_cfm[B2E_MAX_BOXES];
  int self_a[B2E_MAX_SELF_PAIRS], self_b[B2E_MAX_SELF_PAIRS];
  int sph_flags[B2E_MAX_SPHERES];
  int n_caps, cap_link[B2E_MAX_CAPS];
  float cap_p0[B2E_MAX_CAPS][3], cap_p1[B2E_MAX_CAPS][3], cap_r[B2E_MAX_CAPS], cap_mu[B2E_MAX_CAPS];
};

// warp-uniform part of the model: passed by value (constant bank), no loads on the critical path
struct DevModelU {
  int n_links, n_dof, ee_link, n_spheres, fk_rounds, acc_rounds, n_boxes, n_self_pairs, n_caps;
  unsigned ee_dofmask;
  float base_pos[3], base_rot[9], ee_com[3];
  int parent[TLMAX];
};

struct DevState {
  int B;
  float* q; float* qd; float* obj_pose; float* obj_vel; float* target; float* mtarget;
  int* counters; int* cache_key; float* cache_lam; float* hand_pose; int* status; float* raw_obs; float* contacts;
  float* scratch;  // [B][SCRATCH_PER_ENV]: A | W | W^T of a big system when no overflow slot of the block is free
  float* shaping;
};

struct b2e_sim {
  int B, device;
  int tree;   // 0: 16-lane group kernel (<= 12 links, <= 9 dofs: the Panda), 1: warp-per-env tree kernel (<= 32 bodies: the iCub)
  b2e_model model;
  b2e_params params;
  DevModel* d_model;
  DevModelU umodel;
  DevState st;
  void* fields[B2E_F_COUNT];
  float *d_action, *d_obs, *d_reward, *d_done;     // staging for the host-buffer entry point
  float *h_action, *h_obs, *h_reward, *h_done;     // pinned
  int64_t launches;
  int record_contacts;
  cudaEvent_t ev0, ev1;
  cudaStream_t pstream[2];   // chunk pipeline of the page-locked host path
  int* d_sched;              // cost-ordered scheduling state (see NBK_MAIN)
  int sched_seq;
  int tail_wpb;              // warps per block of the tail launch
  int tail_blocks;           // blocks of the tail launch (capacity = blocks x warps x 2 environments, <= TAIL_CAP)
  int tail_excl;             // tail blocks ask for enough shared memory to have their SM to themselves
  int tail_cost;             // estimated solve cycles from which an environment is stepped by the tail launch
  cudaStream_t tstream;      // the tail launch's stream (high priority)
  cudaEvent_t ev_fork, ev_join;
  cudaEvent_t ev_order;      // recorded after every launch: the next entry point's stream waits on it, so calls on
                             // DIFFERENT streams (a torch side stream, the library's own copy streams) stay ordered
};

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, const char* detail) {
  snprintf(g_err, sizeof(g_err), fmt, detail ? detail : "");
  return code;
}
#define CUDA_TRY(x)                                                            \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) return fail(B2E_ECUDA, #x ": %s", cudaGetErrorString(e_)); \
  } while (0)

// Entry points are ordered with respect to each other whatever stream they are given: every launch is followed by
// order_end (event record on its stream), every entry point starts with order_begin (its stream waits on that event).
// On one and the same stream both are no-ops for the GPU.
#ifdef B2E_EMU
#define ORDER_BEGIN(s, stream) do { } while (0)
#define ORDER_END(s, stream) do { } while (0)
#else
#define ORDER_BEGIN(s, stream) CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)(stream), (s)->ev_order, 0))
#define ORDER_END(s, stream) CUDA_TRY(cudaEventRecord((s)->ev_order, (cudaStream_t)(stream)))
#endif

// ------------------------------------------------------------------------------------------
// small device math (all on register arrays with static indexing)

__device__ __forceinline__ void m3mul(const float* a, const float* b, float* c) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
__device__ __forceinline__ void m3vec(const float* a, const float* v, float* o) {
  float x = a[0] * v[0] + a[1] * v[1] + a[2] * v[2];
  float y = a[3] * v[0] + a[4] * v[1] + a[5] * v[2];
  float z = a[6] * v[0] + a[7] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void m3tvec(const float* a, const float* v, float* o) {
  float x = a[0] * v[0] + a[3] * v[1] + a[6] * v[2];
  float y = a[1] * v[0] + a[4] * v[1] + a[7] * v[2];
  float z = a[2] * v[0] + a[5] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
  float x = a[1] * b[2] - a[2] * b[1];
  float y = a[2] * b[0] - a[0] * b[2];
  float z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ __forceinline__ void quat_to_mat(const float* q, float* R) {
  float x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ void mat_to_quat(const float* R, float* q) {
  float tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    float s = sqrtf(tr + 1) * 2;
    q[3] = s / 4; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    float s = sqrtf(1 + R[0] - R[4] - R[8]) * 2;
    q[3] = (R[7] - R[5]) / s; q[0] = s / 4; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    float s = sqrtf(1 + R[4] - R[0] - R[8]) * 2;
    q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = s / 4; q[2] = (R[5] + R[7]) / s;
  } else {
    float s = sqrtf(1 + R[8] - R[0] - R[4]) * 2;
    q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = s / 4;
  }
}
__device__ __forceinline__ void quat_mul(const float* a, const float* b, float* o) {
  float x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  float y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  float z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  float w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
// p.getEulerFromQuaternion restated (panda_env.py:160, world_env.py:119) [EXT-recalled].  Not inlined:
// three call sites with large libm bodies; arguments and result travel in registers.
__device__ __noinline__ float3 quat_to_euler3(float x, float y, float z, float w) {
  float sqx = x * x, sqy = y * y, sqz = z * z, sqw = w * w;
  float sarg = -2 * (x * z - w * y);
  float3 e;
  if (sarg <= -0.99999f) {
    e.x = 0; e.y = -1.57079632679489661923f; e.z = 2 * atan2f(x, -y);
  } else if (sarg >= 0.99999f) {
    e.x = 0; e.y = 1.57079632679489661923f; e.z = 2 * atan2f(-x, y);
  } else {
    e.x = atan2f(2 * (y * z + w * x), sqw - sqx - sqy + sqz);
    e.y = asinf(sarg);
    e.z = atan2f(2 * (x * y + w * z), sqw + sqx - sqy - sqz);
  }
  return e;
}
__device__ __forceinline__ void quat_to_euler(const float* q, float* e) {
  const float3 r = quat_to_euler3(q[0], q[1], q[2], q[3]);
  e[0] = r.x; e[1] = r.y; e[2] = r.z;
}
// p.getQuaternionFromEuler restated (panda_push_gym_env.py:169,172) [EXT-recalled]
__device__ __forceinline__ void euler_to_quat(const float* e, float* q) {
  float sr, cr, sp, cp, sy, cy;
  sincosf(e[0] * 0.5f, &sr, &cr);
  sincosf(e[1] * 0.5f, &sp, &cp);
  sincosf(e[2] * 0.5f, &sy, &cy);
  q[0] = sr * cp * cy - cr * sp * sy;
  q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy;
  q[3] = cr * cp * cy + sr * sp * sy;
}
__device__ __forceinline__ void plane_space(const float* n, float* p, float* q) {
  if (fabsf(n[2]) > 0.7071067811865475244f) {
    float a = n[1] * n[1] + n[2] * n[2];
    float k = 1.0f / sqrtf(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    float a = n[0] * n[0] + n[1] * n[1];
    float k = 1.0f / sqrtf(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

// ------------------------------------------------------------------------------------------
// Environment groups: TWO environments per warp, 16 lanes each (every stage of the pipeline needs
// <= 16 lanes: 12 links, 9 dofs, 8 cube vertices, 14 spheres, 15 velocity components, 16 rows per set).
// One instruction serves both environments.
#define GL 16
struct Grp {
  unsigned hm;  // lanes of this group inside the warp (for masking warp-wide votes)
  int sh;       // bit offset of the group inside the warp (0 or 16)
  int lane;     // lane inside the group, 0..15
};
// The two groups of a warp execute every collective TOGETHER (constant full member mask, 16-wide shuffle
// segments): control flow around collectives is kept warp-uniform — loops run for the longer of the two
// groups, a group that is done keeps pace with frozen state — so no convergence checks are generated.
#define SHF(v, src) __shfl_sync(FULL, (v), (src), GL)
__device__ __forceinline__ unsigned gballot(const Grp& g, bool p) { return (__ballot_sync(FULL, p) >> g.sh) & 0xffffu; }
__device__ __forceinline__ bool gany(const Grp& g, bool p) { return (__ballot_sync(FULL, p) & g.hm) != 0u; }
__device__ __forceinline__ float gmaxf(const Grp& g, float v) {  // max of non-negative floats over the group
  const unsigned b = __float_as_uint(v);
  const unsigned lo = __reduce_max_sync(FULL, g.sh ? 0u : b), hi = __reduce_max_sync(FULL, g.sh ? b : 0u);
  return __uint_as_float(g.sh ? hi : lo);
}
__device__ __forceinline__ void gsync(const Grp& g) { (void)g; __syncwarp(); }

// per-environment shared memory (4.6 KB)
struct Contact {   // 16 words
  int key, type, link, link2;
  float pA[3], pB[3], n[3];
  float dist, mu, erp;
};
struct EnvSmem {
  float A[GL * GL];            // generic-row block of the Delassus matrix, A[c*16 + r], when <= 16 generic rows
                               // (else: per-env global scratch, stride GMAX)
  float W[GL * WSTRIDE];       // W[g][k] = (M^-1 J_g^T)_k of generic row g; k<9 arm dofs, 9..14 cube (lin, ang)
  float WT[NDMAX * GL];        // WT[d][g] = W[g][d]: what a motor row reads (unit stride over g).  A 16-row system
                               // solved in a 2/3-set instantiation (its warp mate is big) reads up to 31 floats past
                               // the end for its non-existent rows: T follows, which nobody writes during the solve
  float T[TLMAX][12];          // link world transforms: R (9) + p (3); dead after collision detection: pgs_solve stages the
                               // impulses of the generic rows here for the friction bounds (3 * GL <= TLMAX * 12 floats)
  float S[NDMAX][6];           // world spatial axes about O=base: (w, v_O)
  float Minv[NDMAX][NDMAX + 1];
  float vstar[16];             // unconstrained velocities (9 arm + 6 cube)
  float mlam[16];              // motor-row impulses (lane = dof)
  float glam[GMAX];            // generic-row impulses
  Contact con[B2E_MAX_CONTACTS];
  float con_cfm[B2E_MAX_CONTACTS];
  int lim_d[4];                // limit rows: dof | side << 8
  float lim_dist[4];
  float cost[4];               // [0]: estimated cycles of this environment's last solve (scheduling class)
  int ckey[B2E_CACHE_SLOTS];
  float clam[B2E_CACHE_SLOTS][3];
};

// Storage of one big constraint system (17..39 generic rows): block-level overflow slot in shared memory,
// or — when the block's slots are taken — the environment's global scratch (same layout).
static_assert(offsetof(EnvSmem, T) % 16 == 0 && sizeof(EnvSmem) % 16 == 0, "EnvSmem: A | W | WT are zeroed with 16-byte stores");
struct BigSlot {
  float A[BIGS * BIGS];      // generic x generic Delassus block, A[c*BIGS + r]
  float W[BIGS * WSTRIDE];   // W[g][k]
  float WT[NDMAX * BIGS + 8]; // WT[d][g] = W[g][d]: motor-row coupling read with unit stride over g (+8: lanes of the
                             // third row set run to g = 47)
};

static_assert(3 * GL <= TLMAX * 12, "pgs_solve stages the generic-row impulses in EnvSmem::T");
static_assert(sizeof(BigSlot) % 16 == 0, "BigSlot is zeroed with 16-byte stores");
// ------------------------------------------------------------------------------------------
// forward kinematics: lane = link.  Composition along the tree by pointer jumping.
__device__ __forceinline__ void fk_lanes(const DevModel* __restrict__ M, const DevModelU& U, const Grp& g, float qi,
                                         float* R, float* p) {
  const int lane = g.lane;
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
    float t = 1 - c, x = ax[0], y = ax[1], z = ax[2];
    float Rq[9] = {t * x * x + c,     t * x * y - s * z, t * x * z + s * y,
                   t * x * y + s * z, t * y * y + c,     t * y * z - s * x,
                   t * x * z - s * y, t * y * z + s * x, t * z * z + c};
    m3mul(jr, Rq, R);
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = jr[k];
    if (jt == B2E_JOINT_PRISMATIC) {
      float t[3] = {ax[0] * qi, ax[1] * qi, ax[2] * qi}, o[3];
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
    for (int k = 0; k < 9; k++) Ra[k] = SHF(R[k], src);
#pragma unroll
    for (int k = 0; k < 3; k++) pa[k] = SHF(p[k], src);
    const int anca = SHF(anc, src);
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

// inclusive sum over the path base..link of a 6-vector held per link lane (pointer jumping)
__device__ __forceinline__ void path_sum6(const DevModel* __restrict__ M, const DevModelU& U, const Grp& g, float* x) {
  int anc = g.lane < U.n_links ? __ldg(&M->parent[g.lane]) : -1;
  const int rounds = U.fk_rounds;
  for (int rd = 0; rd < rounds; rd++) {
    const int src = anc < 0 ? 0 : anc;
    float xa[6];
#pragma unroll
    for (int k = 0; k < 6; k++) xa[k] = SHF(x[k], src);
    const int anca = SHF(anc, src);
    if (anc >= 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) x[k] += xa[k];
      anc = anca;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Damped-least-squares IK (p.calculateInverseKinematics, panda_env.py:269-272): same statement as
// oracle/b2oracle.c ik_dls.  lane = link for FK + Jacobian columns, lane = row (<6) for the 6x6
// system (J J^T + lambda I) x = e solved by Gauss-Jordan over shuffles, lane = dof for dq = J^T x.
// `scr` is 6*16 floats of group-private shared memory.  Returns the joint target of this dof lane.
__device__ __noinline__ float ik_solve(float* scr, const DevModel* __restrict__ M, const DevModelU& U, int max_iters,
                                       float residual, float damping, unsigned hm, int sh, int lane_, float my_q,
                                       float tpx, float tpy, float tpz, float tqx, float tqy, float tqz, float tqw) {
  const Grp g = {hm, sh, lane_};
  const int lane = lane_;
  const int nl = U.n_links, nd = U.n_dof, ee = U.ee_link;
  const int li = lane < nl ? lane : 0;
  const int my_dof = __ldg(&M->dof[li]);
  const bool has_dof = lane < nl && my_dof >= 0 && ((U.ee_dofmask >> (my_dof < 0 ? 0 : my_dof)) & 1);
  const int jt = __ldg(&M->jtype[li]);
  const float ax[3] = {__ldg(&M->axis[li][0]), __ldg(&M->axis[li][1]), __ldg(&M->axis[li][2])};
  float qv = my_q;
  bool fin = false;   // this group reached the residual (the loop is shared by both groups of the warp)
  for (int it = 0; it < max_iters; it++) {
    LANE_SITE(2);
    float R[9], p[3];
    const float qi = SHF(qv, my_dof < 0 ? 0 : my_dof);
    fk_lanes(M, U, g, (lane < nl && my_dof >= 0) ? qi : 0.f, R, p);
    float pe[3], Re[9];
#pragma unroll
    for (int k = 0; k < 3; k++) pe[k] = SHF(p[k], ee);
#pragma unroll
    for (int k = 0; k < 9; k++) Re[k] = SHF(R[k], ee);
    const float dp[3] = {tpx - pe[0], tpy - pe[1], tpz - pe[2]};
    if (sqrtf(dot3(dp, dp)) <= residual) fin = true;
    if (__all_sync(FULL, fin)) break;
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
    // Jacobian column of this link's joint
    float c[6] = {0, 0, 0, 0, 0, 0};
    if (has_dof) {
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
    gsync(g);
    for (int k = lane; k < 6 * 16; k += GL) scr[k] = 0.f;
    gsync(g);
    if (lane < nl && my_dof >= 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) scr[k * 16 + my_dof] = c[k];
    }
    gsync(g);
    // lane r < 6: row r of J J^T + lambda I, augmented with e_r
    const int r = lane < 6 ? lane : 0;
    float Ur[7];
#pragma unroll
    for (int cc2 = 0; cc2 < 6; cc2++) {
      float acc = (cc2 == r) ? damping : 0.f;
      for (int d = 0; d < nd; d++) acc = fmaf(scr[r * 16 + d], scr[cc2 * 16 + d], acc);
      Ur[cc2] = acc;
    }
    Ur[6] = r < 3 ? (r == 0 ? dp[0] : (r == 1 ? dp[1] : dp[2])) : (r == 3 ? er[0] : (r == 4 ? er[1] : er[2]));
#pragma unroll
    for (int k = 0; k < 6; k++) {
      float rk[7];
#pragma unroll
      for (int j = 0; j < 7; j++) rk[j] = SHF(Ur[j], k);
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
    // dq_d = sum_r J[r][d] x_r  (lane = dof)
    float dq = 0.f;
#pragma unroll
    for (int rr = 0; rr < 6; rr++) {
      const float xr = SHF(Ur[6], rr);
      dq = fmaf(scr[rr * 16 + lane], xr, dq);
    }
    if (lane >= nd) dq = 0.f;
    const float mx = gmaxf(g, fabsf(dq));
    const float scale = mx > 0.78539816339f ? 0.78539816339f / mx : 1.f;
    if (!fin) qv = fmaf(dq, scale, qv);
  }
  gsync(g);
  return qv;
}

// ------------------------------------------------------------------------------------------
// PGS.  Rows of one env: the n_dof position-motor rows (lane = dof, MotorRegs) followed by the
// "generic" rows (limit rows, contact normals, contact frictions) in sets of 16 (lane = row,
// RowRegs<NSG>).  The solver runs on the Delassus form A = J M^-1 J^T; thanks to J_motor = e_d its
// blocks are:  motor x motor = M^-1 (read from sm.Minv),  motor x generic = W_g[d] (read from the W
// table),  generic x generic = A (shared memory when <= 16 generic rows, else global scratch).
// Per-row state: u = rhs - (A lambda)_r is the running velocity error, base = lambda*(1 - cfm*invd),
// so the dependent chain of one row update is FFMA -> FMNMX -> FMNMX -> FADD -> SHFL -> FFMA.
// T = (D+L)^-1 of the motor block M^-1 = L + D + U of the Delassus matrix by forward substitution, lane = row
// (T[k] = T[lane][k]); Ar = row `lane` of M^-1 (identity outside the n_dof block).  Both groups of the warp together.
__device__ __forceinline__ void motor_block_T(const Grp& g, const float* Minv, int nd, float invd, float* Ar, float* T) {
  const int lane = g.lane;
  const bool row = lane < nd;
  const int lc = lane < NDMAX ? lane : NDMAX;
#pragma unroll
  for (int k = 0; k < NDMAX; k++) Ar[k] = (row && k < nd) ? Minv[k * (NDMAX + 1) + lc] : ((k == lane) ? 1.f : 0.f);
  const float idg = row ? invd : 1.f;
#pragma unroll
  for (int k = 0; k < NDMAX; k++) T[k] = (k == lane) ? 1.f : 0.f;
#pragma unroll
  for (int j = 0; j < NDMAX; j++) {
#pragma unroll
    for (int k = 0; k <= j; k++) {
      const float tj = SHF(T[k] * idg, j);       // final row j of T
      if (lane == j) T[k] = tj;
      else if (lane > j) T[k] = fmaf(-Ar[j], tj, T[k]);
    }
  }
}

struct MotorRegs {
  float u, invd, diag, lo, hi, lam, prev;
};
template <int NSG>
struct RowRegs {
  float lam[NSG], u[NSG], gm1[NSG], invd[NSG], diag[NSG], lo[NSG], hi[NSG], mu[NSG], prev[NSG];
  float cc[NSG], lo2[NSG], hi2[NSG];   // per-phase constants of the delta-form row update (rows_prepare)
  int type[NSG], isl[NSG], nidx[NSG];
};

// Row updates are branch-free: the table entries are loaded first (their latency hides under the dependent
// FFMA -> FMNMX -> FMNMX -> FADD -> SHFL chain), and a group that does not visit the row (`active` false)
// broadcasts a zero impulse change; its table entries may be stale but are always finite (step_kernel zeroes the
// tables of the block at its start; the global scratch is zeroed at create), so 0 * entry = 0.
template <int NSG>
__device__ __forceinline__ void motor_step(const Grp& g, MotorRegs& m, RowRegs<NSG>& r, const float* Mrow,
                                           const float* WTr, int i, float lo2, float hi2) {
  // delta form (see generic_step; motor rows have no cfm): dlambda = clamp(u invd, lo - lambda, hi - lambda) with the two
  // bounds prepared once per sweep by the caller (a row is visited once per sweep); a group that does not sweep its motor
  // rows in this pass has lo2 = hi2 = 0 and broadcasts 0
  // Mrow / WTr: this lane's column of row i of M^-1 (padded to NDMAX + 1, pad = 0) and of the transposed W table
  const float cm = Mrow[0];
  float cg[NSG];   // A[generic g][motor i] = W_g[i], read from the transposed table (unit stride over g)
#pragma unroll
  for (int s = 0; s < NSG; s++) cg[s] = WTr[GL * s];
  const float dl = fminf(fmaxf(m.u * m.invd, lo2), hi2);
  const float dli = SHF(dl, i);
  m.u = fmaf(-cm, dli, m.u);                    // inactive: dli = 0 and the tables are finite (zeroed at kernel start)
  if (g.lane == i) m.lam += dl;
#pragma unroll
  for (int s = 0; s < NSG; s++) r.u[s] = fmaf(-cg[s], dli, r.u[s]);
}

// Delta form of the row update.  With c = lambda (gg - 1), lo' = lo - lambda, hi' = hi - lambda prepared once per sweep
// phase (a row is visited once per phase, so its lambda is constant until then):
//     dlambda = clamp(u invd + c, lo', hi')        [ = clamp(u invd + lambda gg, lo, hi) - lambda ]
// which leaves FFMA -> FMNMX -> FMNMX -> SHFL -> FFMA on the dependent chain (tools/micro/row_chain2.cu: 50 instead of 68
// cycles per row for a lone warp).  A row this group does not visit in the phase has lo' = hi' = 0, i.e. broadcasts 0; every
// table entry is finite (tables are zeroed at kernel start), so no coefficient needs masking.
template <int NSG>
__device__ __forceinline__ void rows_prepare(const Grp& g, RowRegs<NSG>& r, unsigned m0, unsigned m1, unsigned m2) {
#pragma unroll
  for (int s = 0; s < NSG; s++) {
    const unsigned mk = s == 0 ? m0 : (s == 1 ? m1 : m2);
    const bool act = (mk >> g.lane) & 1u;
    r.cc[s] = r.lam[s] * r.gm1[s];
    r.lo2[s] = act ? r.lo[s] - r.lam[s] : 0.f;
    r.hi2[s] = act ? r.hi[s] - r.lam[s] : 0.f;
  }
}

template <int NSG, int SI>
__device__ __forceinline__ void generic_step(const Grp& g, MotorRegs& m, RowRegs<NSG>& r, const float* Arow,
                                             const float* Wrow, int li) {
  // Arow / Wrow already point at this lane's column of the row
  float ca[NSG];
#pragma unroll
  for (int s = 0; s < NSG; s++) ca[s] = Arow[GL * s];
  const float cw = Wrow[0];
  const float dl = fminf(fmaxf(fmaf(r.u[SI], r.invd[SI], r.cc[SI]), r.lo2[SI]), r.hi2[SI]);
  const float dli = SHF(dl, li);
  r.u[SI] = fmaf(-ca[SI], dli, r.u[SI]);        // the running error the next rows of this set wait for: first
  if (g.lane == li) r.lam[SI] += dl;
  m.u = fmaf(-cw, dli, m.u);                    // m.u of a finished arm island is dead
#pragma unroll
  for (int s = 0; s < NSG; s++)
    if (s != SI) r.u[s] = fmaf(-ca[s], dli, r.u[s]);
}

// Visit the rows of one 16-row set whose bits are set in `mk`, in ascending order.  The loop is a counted loop
// over the index range spanned by BOTH groups' masks (warp-uniform trip count, no find-first-set chain; the
// masks are contiguous ranges in practice: limits, then normals, then frictions, cube-table contacts first);
// a group skips (dl = 0, no table access) the rows it does not visit itself.
template <int NSG, int SI>
__device__ __forceinline__ void sweep_set(const Grp& g, MotorRegs& m, RowRegs<NSG>& r, const float* A, const float* W,
                                          int AS, unsigned w) {
  // w: union of the two groups' row masks of this set (__reduce_or_sync: provably warp-uniform trip count)
  if (w == 0u) return;
  const int lo = __ffs(w) - 1, hi = 32 - __clz(w);
  const float* Arow = A + (GL * SI + lo) * AS + g.lane;
  const float* Wrow = W + (GL * SI + lo) * WSTRIDE + g.lane;
  B2E_UNROLL(SWEEP_UNROLL)
  for (int i = lo; i < hi; i++) {
    generic_step<NSG, SI>(g, m, r, Arow, Wrow, i);
    Arow += AS;
    Wrow += WSTRIDE;
  }
}

template <int NSG>
__device__ __forceinline__ void sweep_generic(const Grp& g, MotorRegs& m, RowRegs<NSG>& r, const float* A,
                                              const float* W, int AS, unsigned m0, unsigned m1, unsigned m2, unsigned w0,
                                              unsigned w1, unsigned w2) {
  // m*: rows of each set THIS group visits in the phase; w*: their unions over the two groups of the warp (loop bounds)
  constexpr int S1 = NSG > 1 ? 1 : 0, S2 = NSG > 2 ? 2 : 0;
  rows_prepare<NSG>(g, r, m0, m1, m2);
  sweep_set<NSG, 0>(g, m, r, A, W, AS, w0);
  if (NSG > 1) sweep_set<NSG, S1>(g, m, r, A, W, AS, w1);
  if (NSG > 2) sweep_set<NSG, S2>(g, m, r, A, W, AS, w2);
}

template <int NSG>
__device__ __forceinline__ int pgs_solve(const Grp& g, MotorRegs& m, RowRegs<NSG>& r, const float* A, const float* W,
                                         const float* WT, int AS, const float* Minv, float* stage, int nd, int RG, int fric_start,
                                         bool coupled, bool has_cube_rows, bool arm_sweep, int max_iters, float tol,
                                         float* prof = nullptr) {   // prof (-DPROFILE_SWEEP builds): cycles per part of the sweeps
#ifdef PROFILE_SWEEP
  float pacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#define PSW(k) { const long long t_ = clock64(); pacc[k] += (float)(t_ - t_psw); t_psw = t_; }
#else
#define PSW(k) do { } while (0)
#endif
  // The motor rows of a coupled / limit-carrying arm island are swept serially, in delta form (motor_step).  Round 1 solved them
  // as one block per sweep (dlambda = (D+L)^-1 u, then u -= M^-1 dlambda: two 9-wide mat-vecs over shuffles, falling back to
  // the serial rows when a bound would activate); measured in round 2 with cycle stamps inside the sweep (tools/stage_profile.py,
  // -DPROFILE_SWEEP): for the lone warp of a sweep-capped solve the block form cost ~1700 cycles per sweep against ~650 for the
  // nine serial rows (ptxas interleaves its shuffles with their consumers on reused registers, so they complete one by one),
  // and whole-batch throughput went 36.6 M -> 40.4 M env-steps/s with the serial rows alone.  `stage` (the dead link-transform
  // table) stages the impulses for the friction bounds.
  // generic-row masks per set: island (arm / cube) x phase (non-friction, friction)
  unsigned arm_nf[3] = {0, 0, 0}, arm_f[3] = {0, 0, 0}, cube_nf[3] = {0, 0, 0}, cube_f[3] = {0, 0, 0};
#pragma unroll
  for (int s = 0; s < NSG; s++) {
    const int gi = GL * s + g.lane;
    const bool valid = gi < RG, fr = gi >= fric_start;
    const bool cube = !coupled && r.isl[s] == 1;
    arm_nf[s] = gballot(g, valid && !cube && !fr);
    arm_f[s] = gballot(g, valid && !cube && fr);
    cube_nf[s] = gballot(g, valid && cube && !fr);
    cube_f[s] = gballot(g, valid && cube && fr);
  }
  // Control state of the shared sweep loop is kept warp-uniform: "island finished" flags of BOTH groups (A = lanes 0-15,
  // B = lanes 16-31), maintained from two ballots per sweep; the unions of the row masks (loop bounds) are recomputed only
  // when a flag flips.
  bool done0 = !arm_sweep, done1 = !has_cube_rows || coupled;
  bool dA0, dB0, dA1, dB1;
  {
    const unsigned b0 = __ballot_sync(FULL, done0), b1 = __ballot_sync(FULL, done1);
    dA0 = b0 & 1u; dB0 = (b0 >> GL) & 1u; dA1 = b1 & 1u; dB1 = (b1 >> GL) & 1u;
  }
  int my_it = (done0 && done1) ? 0 : -1;   // sweep count of this group (the loop itself is shared by the warp)
  unsigned wn[3] = {0, 0, 0}, wf[3] = {0, 0, 0};
  bool fresh = true;
  for (int it = 0; it < max_iters; it++) {
    LANE_SITE(0);
    if (dA0 && dB0 && dA1 && dB1) break;
    const unsigned a0 = done0 ? 0u : 0xffffffffu, c0 = done1 ? 0u : 0xffffffffu;
    if (fresh) {
#pragma unroll
      for (int s = 0; s < NSG; s++) {
        wn[s] = __reduce_or_sync(FULL, (arm_nf[s] & a0) | (cube_nf[s] & c0));
        wf[s] = __reduce_or_sync(FULL, (arm_f[s] & a0) | (cube_f[s] & c0));
      }
      fresh = false;
    }
#ifdef PROFILE_SWEEP
    long long t_psw = clock64();
#endif
    m.prev = m.lam;
#pragma unroll
    for (int s = 0; s < NSG; s++) r.prev[s] = r.lam[s];
    if (!(dA0 && dB0)) {  // motor rows first (Bullet: non-contact constraints, then normals, then frictions)
      PSW(5);   // (profiling builds) previous-impulse copies
      const float mlo2 = done0 ? 0.f : m.lo - m.lam, mhi2 = done0 ? 0.f : m.hi - m.lam;
      const float* Mrow = Minv + (g.lane < NDMAX ? g.lane : NDMAX);
      const float* WTr = WT + g.lane;
      B2E_UNROLL(SWEEP_UNROLL)
      for (int i = 0; i < nd; i++) {
        motor_step<NSG>(g, m, r, Mrow, WTr, i, mlo2, mhi2);
        Mrow += NDMAX + 1;
        WTr += AS;
      }
    }
    PSW(0);   // masks + motor rows
    sweep_generic<NSG>(g, m, r, A, W, AS, (arm_nf[0] & a0) | (cube_nf[0] & c0), (arm_nf[1] & a0) | (cube_nf[1] & c0),
                       (arm_nf[2] & a0) | (cube_nf[2] & c0), wn[0], wn[1], wn[2]);
    PSW(1);   // non-friction rows
    if ((wf[0] | wf[1] | wf[2]) != 0u) {
      // friction bounds from the current normal impulses (mu * lambda_n): every lane stages its impulses, a friction row reads
      // its contact's normal impulse back (3 stores + 3 loads instead of NSG^2 shuffles with their selects)
#pragma unroll
      for (int s = 0; s < NSG; s++) stage[GL * s + g.lane] = r.lam[s];
      gsync(g);
#pragma unroll
      for (int s = 0; s < NSG; s++) {
        const float v = stage[r.nidx[s]];
        if (r.type[s] == ROW_FRICTION) {
          const float lim = r.mu[s] * v;
          r.lo[s] = -lim; r.hi[s] = lim;
        }
      }
      gsync(g);
      PSW(2);   // friction bounds
      sweep_generic<NSG>(g, m, r, A, W, AS, (arm_f[0] & a0) | (cube_f[0] & c0), (arm_f[1] & a0) | (cube_f[1] & c0),
                         (arm_f[2] & a0) | (cube_f[2] & c0), wf[0], wf[1], wf[2]);
      PSW(3);   // friction rows
    }
    float ra = 0.f, rc = 0.f;
    if (!done0 && g.lane < nd) {
      const float rv = (m.lam - m.prev) * m.diag;
      ra = rv * rv;
    }
#pragma unroll
    for (int s = 0; s < NSG; s++) {
      float rv = (r.lam[s] - r.prev[s]) * r.diag[s];  // impulse change of this sweep
      rv = rv * rv;
      if (GL * s + g.lane < RG) {
        if (coupled || r.isl[s] == 0) ra = fmaxf(ra, rv); else rc = fmaxf(rc, rv);
      }
    }
    // an island is finished when the largest squared velocity change of the sweep is <= tol: "no lane above tol"
    const unsigned ba = __ballot_sync(FULL, !(ra <= tol)), bc = __ballot_sync(FULL, !(rc <= tol));
    const bool nA0 = dA0 || !(ba & 0xffffu), nB0 = dB0 || !(ba >> GL), nA1 = dA1 || !(bc & 0xffffu), nB1 = dB1 || !(bc >> GL);
    fresh = (nA0 != dA0) || (nB0 != dB0) || (nA1 != dA1) || (nB1 != dB1);
    dA0 = nA0; dB0 = nB0; dA1 = nA1; dB1 = nB1;
    done0 = g.sh ? dB0 : dA0;
    done1 = g.sh ? dB1 : dA1;
    if (my_it < 0 && done0 && done1) my_it = it + 1;
    PSW(4);   // residual + control
  }
#ifdef PROFILE_SWEEP
  if (prof && g.lane == 0)
    for (int k = 0; k < 6; k++) prof[k] = pacc[k];
#endif
#undef PSW
  return my_it < 0 ? max_iters : my_it;
}

// Arm island made only of the n_dof position-motor rows (no limit row, no arm contact): bounds are
// +-max_force*dt and in practice never active, so one Gauss-Seidel sweep over those rows IS the affine
// map lambda' = G lambda + c with G = -(D+L)^-1 U, c = (D+L)^-1 b (L + D + U = M^-1, the motor block of
// the Delassus matrix).  Same iterates in exact arithmetic, same per-sweep residual test; the nine
// serial, shuffle-dependent row updates of a sweep become one 9-wide mat-vec.  If a bound would
// activate the caller falls back to the serial sweep.  Returns the sweep count, or -1 on fallback.
__device__ __forceinline__ int arm_affine_solve(const Grp& g, const float* Minv, int nd, float b, float invd, float diag,
                                                float lo, float hi, int max_iters, float tol, float& lam_out) {
  const int lane = g.lane;
  const bool row = lane < nd;
  float Ar[NDMAX], T[NDMAX];
  motor_block_T(g, Minv, nd, invd, Ar, T);
  // G = -T U, c = T b
  float G[NDMAX];
#pragma unroll
  for (int k = 0; k < NDMAX; k++) G[k] = 0.f;
  float c = 0.f;
  const float bb = row ? b : 0.f;
#pragma unroll
  for (int j = 0; j < NDMAX; j++) {
    c = fmaf(T[j], SHF(bb, j), c);
#pragma unroll
    for (int k = j + 1; k < NDMAX; k++) G[k] = fmaf(-T[j], SHF(Ar[k], j), G[k]);
  }
  float lam = 0.f;
  int my_it = -1;          // >= 0: converged after that many sweeps; -2: a bound would activate (fallback)
  bool openA = true, openB = true;   // warp-uniform: group A / B still iterating (two ballots per sweep)
  for (int it = 0; it < max_iters; it++) {
    LANE_SITE(1);
    if (!openA && !openB) break;   // the loop is shared by both groups of the warp
    float a0 = c, a1 = 0.f;
#pragma unroll
    for (int k = 1; k < NDMAX; k += 2) a0 = fmaf(G[k], SHF(lam, k), a0);
#pragma unroll
    for (int k = 2; k < NDMAX; k += 2) a1 = fmaf(G[k], SHF(lam, k), a1);
    const float nl = a0 + a1;
    const float rv = row ? (nl - lam) * diag : 0.f;
    const unsigned bcl = __ballot_sync(FULL, row && !(nl >= lo && nl <= hi)), bres = __ballot_sync(FULL, !(rv * rv <= tol));
    const bool clamp = (bcl & g.hm) != 0u, conv = (bres & g.hm) == 0u;
    if (my_it == -1) {
      if (clamp) my_it = -2;
      else {
        lam = nl;
        if (conv) my_it = it + 1;
      }
    }
    openA = openA && !(bcl & 0xffffu) && (bres & 0xffffu);
    openB = openB && !(bcl >> GL) && (bres >> GL);
  }
  if (my_it == -2) return -1;
  lam_out = lam;
  return my_it < 0 ? max_iters : my_it;
}

// Build the motor row of this dof lane and the generic row(s) of this lane, the W table, the generic
// block of the Delassus matrix, warm start, solve, and leave the impulses in sm.mlam[] / sm.glam[].
// Returns the PGS iteration count.
template <int NSG, bool GB>   // GB: some group of the warp keeps its big system in GLOBAL scratch (no slot was free)
__device__ __noinline__ int build_and_solve(EnvSmem& sm, const DevModel* __restrict__ M, const DevModelU& U,
                                            const b2e_params& P, unsigned hm, int sh, int lane, int nd, int nlim, int nc,
                                            float my_q, float my_target, float my_kp, float my_mv, float cpx, float cpy, float cpz,
                                            int big_off, float* gscratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Grp g = {hm, sh, lane};
  const float cpos[3] = {cpx, cpy, cpz};
  // up to 16 generic rows live in the environment's own shared memory; a bigger system uses `big`
  // (an overflow slot of the block, or the env's global scratch), which also carries W^T
  const bool use_big = (NSG > 1) && (nlim + 3 * nc > GL);
  // row stride; passed through a shuffle (identity: source lane = own lane) so that ptxas keeps it in a register — it
  // otherwise re-selects between the two constants in every row-loop iteration (three dependent instructions on the pointer)
  const int AS = SHF(use_big ? BIGS : GL, lane);
  // the slot address is formed here from the shared-memory base so that, in the GB = false instantiations,
  // every table access is a shared-memory access (no generic loads in the row loops)
  float* big = reinterpret_cast<float*>(smem_raw + (big_off < 0 ? 0 : big_off));   // byte offset of the claimed overflow slot
  if (GB && big_off < 0) big = gscratch;
  float* A = use_big ? big : sm.A;
  float* W = use_big ? big + BIGS * BIGS : sm.W;
  float* WT = use_big ? big + BIGS * BIGS + BIGS * WSTRIDE : sm.WT;
  const float* Minv = &sm.Minv[0][0];
  const int fric_start = nlim + nc;   // generic index of the first friction row
  const int RG = nlim + 3 * nc;       // generic rows
  const int RGw = (int)__reduce_max_sync(FULL, (unsigned)RG);   // larger count of the two groups: shared loop bounds (uniform)
  const float dt = P.dt, inv_dt = 1.0f / P.dt;
  const float cinv_m = 1.0f / P.cube_mass, cinv_I = 1.0f / P.cube_inertia;

#ifdef PROFILE_STAGES
  const long long t_bs0 = clock64();
#ifdef PROFILE_WARM   // probe build (with -DPROFILE_STAGES): the three solve-part outputs carry ACTIVE-LANE counts instead of cycles,
  // cost[1] = entry + 100 x after rows/W + 10000 x after the Delassus block, cost[2] = after the cache look-up / the warm-start
  // apply / the sweeps, cost[3] = (kernel body) after the collision stage / before / after the overflow-slot claim.  32 everywhere
  // = the warp is converged; 31 = a straggler lane, every shuffle takes the divergent path (WARPSYNC.COLLECTIVE)
  int am0 = __popc(__activemask()), am1 = 0, am2 = 0, am3 = 0, am4 = 0, am5 = 0;
#endif
#endif
  // ---- motor row of dof `lane` (btMultiBodyJointMotor: velocity target kp*(target-q)/dt, kd = 1, erp = 1) ----
  MotorRegs m;
  {
    const bool isd = lane < nd;
    const int d = isd ? lane : 0;
    float desired = my_kp * (my_target - my_q) * inv_dt;
    const float mv = my_mv;
    if (mv > 0) desired = fminf(fmaxf(desired, -mv), mv);
    const float diag = isd ? sm.Minv[d][d] : 1.f;
    m.diag = diag;
    m.invd = isd ? 1.0f / diag : 0.f;
    m.u = isd ? desired - sm.vstar[d] : 0.f;
    m.hi = isd ? __ldg(&M->max_force[d]) * dt : 0.f;
    m.lo = -m.hi;
    m.lam = 0.f;
    m.prev = 0.f;
  }

  // ---- generic rows ----
  RowRegs<NSG> rr;
  float Jall[NSG][15];
  bool coupled = false, has_cube = false, arm_generic = false;
#pragma unroll
  for (int s = 0; s < NSG; s++) {
    const int gi = GL * s + lane;
    const bool valid = gi < RG;
    float* J = Jall[s];
#pragma unroll
    for (int k = 0; k < 15; k++) J[k] = 0.f;
    float desired = 0.f, cfm = 0.f, lo = 0.f, hi = 0.f, mu = 0.f;
    int type = ROW_LIMIT, isl = 0, nidx = 0;
    bool sphere_cube_normal = false;
    if (valid) {
      if (gi < nlim) {  // joint limit row
        type = ROW_LIMIT;
        const int code = sm.lim_d[gi];
        const int d = code & 0xff, side = code >> 8;
#pragma unroll
        for (int k = 0; k < NDMAX; k++) J[k] = (k == d) ? (side ? -1.f : 1.f) : 0.f;
        const float pen = sm.lim_dist[gi] + P.slop;
        desired = pen > 0 ? -pen * inv_dt : -pen * P.erp * inv_dt;
        lo = 0.f; hi = 1e30f;
      } else {  // contact rows
        int c, which;  // which: 0 normal, 1/2 friction
        if (gi < fric_start) { c = gi - nlim; which = 0; }
        else { c = (gi - fric_start) >> 1; which = 1 + ((gi - fric_start) & 1); }
        const Contact& ct = sm.con[c];
        float n[3] = {ct.n[0], ct.n[1], ct.n[2]}, dir[3];
        if (which == 0) { dir[0] = n[0]; dir[1] = n[1]; dir[2] = n[2]; }
        else {
          float t1[3], t2[3];
          plane_space(n, t1, t2);
          if (which == 1) { dir[0] = t1[0]; dir[1] = t1[1]; dir[2] = t1[2]; }
          else { dir[0] = t2[0]; dir[1] = t2[1]; dir[2] = t2[2]; }
        }
        isl = ct.type == CT_CUBE_STATIC ? 1 : 0;
        if (ct.type == CT_CUBE_STATIC) {
          float rel[3] = {ct.pA[0] - cpos[0], ct.pA[1] - cpos[1], ct.pA[2] - cpos[2]}, t[3];
          cross3(rel, dir, t);
#pragma unroll
          for (int k = 0; k < 3; k++) { J[9 + k] = dir[k]; J[12 + k] = t[k]; }
        } else {
          const unsigned mask = __ldg(&M->link_dofmask[ct.link]);
          float rel[3] = {ct.pA[0] - U.base_pos[0], ct.pA[1] - U.base_pos[1], ct.pA[2] - U.base_pos[2]}, wn[3];
          cross3(rel, dir, wn);
#pragma unroll
          for (int d = 0; d < NDMAX; d++) {
            float v = sm.S[d][0] * wn[0] + sm.S[d][1] * wn[1] + sm.S[d][2] * wn[2] + sm.S[d][3] * dir[0] +
                      sm.S[d][4] * dir[1] + sm.S[d][5] * dir[2];
            J[d] = ((mask >> d) & 1) ? v : 0.f;
          }
          if (ct.type == CT_ARM_ARM) {   // self-collision: minus the velocity of the point on link2
            const unsigned mask2 = __ldg(&M->link_dofmask[ct.link2]);
            float rel2[3] = {ct.pB[0] - U.base_pos[0], ct.pB[1] - U.base_pos[1], ct.pB[2] - U.base_pos[2]}, wn2[3];
            cross3(rel2, dir, wn2);
#pragma unroll
            for (int d = 0; d < NDMAX; d++) {
              float v = sm.S[d][0] * wn2[0] + sm.S[d][1] * wn2[1] + sm.S[d][2] * wn2[2] + sm.S[d][3] * dir[0] +
                        sm.S[d][4] * dir[1] + sm.S[d][5] * dir[2];
              J[d] -= ((mask2 >> d) & 1) ? v : 0.f;
            }
          }
          if (ct.type == CT_ARM_CUBE) {
            float relc[3] = {ct.pB[0] - cpos[0], ct.pB[1] - cpos[1], ct.pB[2] - cpos[2]}, t[3];
            cross3(relc, dir, t);
#pragma unroll
            for (int k = 0; k < 3; k++) { J[9 + k] = -dir[k]; J[12 + k] = -t[k]; }
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
    // W = M^-1 J^T.  The generic 9x9 product is only needed when some row of this set has an arm part
    // (limit row or robot contact); rows of cube-table contacts have none.
    float Wv[15];
    const bool arm_part = valid && isl == 0;
    if (gany(g, arm_part)) {
#pragma unroll
      for (int d = 0; d < NDMAX; d++) {
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < NDMAX; e++) acc = fmaf(sm.Minv[d][e], J[e], acc);
        Wv[d] = acc;
      }
    } else {
#pragma unroll
      for (int d = 0; d < NDMAX; d++) Wv[d] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { Wv[9 + k] = J[9 + k] * cinv_m; Wv[12 + k] = J[12 + k] * cinv_I; }
    float diag = 0.f, jv = 0.f;
#pragma unroll
    for (int k = 0; k < 15; k++) { diag = fmaf(J[k], Wv[k], diag); jv = fmaf(J[k], sm.vstar[k], jv); }
    if (valid) {
#pragma unroll
      for (int k = 0; k < 15; k++) W[gi * WSTRIDE + k] = Wv[k];
      W[gi * WSTRIDE + 15] = 0.f;
#pragma unroll
      for (int k = 0; k < NDMAX; k++) WT[k * AS + gi] = Wv[k];
    }
    rr.type[s] = type; rr.isl[s] = isl; rr.nidx[s] = nidx;
    rr.lo[s] = lo; rr.hi[s] = hi; rr.mu[s] = mu;
    rr.diag[s] = valid ? diag + cfm : 0.f;
    rr.invd[s] = valid ? 1.0f / (diag + cfm) : 0.f;
    rr.gm1[s] = -cfm * rr.invd[s];
    rr.u[s] = valid ? desired - jv : 0.f;
    rr.lam[s] = 0.f; rr.prev[s] = 0.f;
    rr.cc[s] = 0.f; rr.lo2[s] = 0.f; rr.hi2[s] = 0.f;
    {  // votes are collectives: evaluate them unconditionally (no short-circuit), then combine
      const bool v1 = gany(g, sphere_cube_normal), v2 = gany(g, valid && isl == 1), v3 = gany(g, arm_part);
      coupled = coupled || v1;
      has_cube = has_cube || v2;
      arm_generic = arm_generic || v3;
    }
  }
  gsync(g);
#if defined(PROFILE_STAGES) && !defined(PROFILE_WARM)
  if (lane == 0) sm.cost[1] = (float)(clock64() - t_bs0);   // rows + W built
#endif
#ifdef PROFILE_WARM
  am1 = __popc(__activemask());
#endif
  // generic x generic block: A[c][r] = J_r . W_c  (symmetric; stored so that row c is contiguous in r)
#pragma unroll
  for (int s = 0; s < NSG; s++) {
    const int gi = GL * s + lane;
    const bool valid = gi < RG;
    const float* J = Jall[s];
    for (int c = 0; c < RG; c++) {
      // row c of W as four 16-byte loads issued together (rows are 64-byte aligned).  With scalar loads placed next to their
      // FFMAs behind the two branches, the 15-term dot product ran load -> FFMA -> load -> FFMA: ~1250 cycles per (set, c) for
      // the lone warp of a big system, 137 k cycles for a 36-row block (cycle stamps, -DPROFILE_STAGES).  Same terms, same order.
      const float4* w4 = reinterpret_cast<const float4*>(W + c * WSTRIDE);
      const float4 wa = w4[0], wb = w4[1], wc = w4[2], wd = w4[3];
      const float w[16] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w, wd.x, wd.y, wd.z, wd.w};
      // branch-free: the arm part of a cube-table column is dropped by a select (its W entries are zero anyway, but J of this
      // lane's row need not be), the cube part of a limit column multiplies zeros (W[c][9..14] = 0 for c < nlim)
      const int cc = c < nlim ? 0 : (c < fric_start ? c - nlim : (c - fric_start) >> 1);
      const bool cube_only = (c >= nlim) && sm.con[cc].type == CT_CUBE_STATIC;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < NDMAX; k++) acc = fmaf(J[k], w[k], acc);
      acc = cube_only ? 0.f : acc;
#pragma unroll
      for (int k = NDMAX; k < 15; k++) acc = fmaf(J[k], w[k], acc);
      if (valid) A[c * AS + gi] = acc;
    }
  }
  gsync(g);
#ifdef PROFILE_WARM
  am2 = __popc(__activemask());
#endif
  // warm start (contact rows): lambda0 = cached impulse * factor, u -= A lambda0
#pragma unroll
  for (int s = 0; s < NSG; s++) {
    const int gi = GL * s + lane;
    if (gi >= nlim && gi < RG) {
      int c, j;
      if (gi < fric_start) { c = gi - nlim; j = 0; }
      else { c = (gi - fric_start) >> 1; j = 1 + ((gi - fric_start) & 1); }
      const int key = sm.con[c].key;
      float l0 = 0.f;
      for (int sl = 0; sl < B2E_CACHE_SLOTS; sl++)
        if (sm.ckey[sl] == key) { l0 = sm.clam[sl][j] * P.warmstart; break; }
      rr.lam[s] = l0;
    }
  }
#ifdef PROFILE_WARM
  am3 = __popc(__activemask());
#endif
  for (int c = 0; c < RGw; c++) {
    float l0 = 0.f;
#pragma unroll
    for (int t = 0; t < NSG; t++) {
      const float vt = SHF(rr.lam[t], c & (GL - 1));
      if ((c >> 4) == t) l0 = vt;
    }
    if (l0 != 0.f && c >= nlim && c < RG) {
#pragma unroll
      for (int s = 0; s < NSG; s++) rr.u[s] = fmaf(-A[c * AS + GL * s + lane], l0, rr.u[s]);
      m.u = fmaf(-W[c * WSTRIDE + lane], l0, m.u);
    }
  }
#ifdef PROFILE_WARM
  am4 = __popc(__activemask());
#elif defined(PROFILE_STAGES)
  if (lane == 0) sm.cost[2] = (float)(clock64() - t_bs0);   // + Delassus block, warm start
#endif
  int iters_arm = -1;
  bool arm_sweep = true;
  const bool arm_simple = !coupled && !arm_generic;
  if (__any_sync(FULL, arm_simple)) {   // both groups run it (shared loop); a non-simple group ignores the result
    float lam_arm = 0.f;
    const int ia = arm_affine_solve(g, Minv, nd, m.u, m.invd, m.diag, m.lo, m.hi, P.solver_iters, P.residual_tol, lam_arm);
    if (arm_simple && ia >= 0) {
      iters_arm = ia;
      if (lane < nd) m.lam = lam_arm;
      arm_sweep = false;
    }
  }
#if defined(PROFILE_STAGES) && !defined(PROFILE_WARM)
  if (lane == 0) sm.cost[3] = (float)(clock64() - t_bs0);   // + affine arm solve
#endif
#ifdef PROFILE_CYCLES
  const long long t_pgs0 = clock64();
#endif
#ifdef PROFILE_SWEEP   // the environment's global scratch is free while its system lives in a slot: its last 8 floats carry the profile
  float* prof = (use_big && big != gscratch) || !use_big ? gscratch + SCRATCH_PER_ENV - 8 : nullptr;
#else
  float* prof = nullptr;
#endif
  int iters = pgs_solve<NSG>(g, m, rr, A, W, WT, AS, Minv, &sm.T[0][0], nd, RG, fric_start, coupled, has_cube, arm_sweep, P.solver_iters,
                            P.residual_tol, prof);
#ifdef PROFILE_CYCLES
  if (lane == 0) sm.lim_dist[3] = (float)(clock64() - t_pgs0);   // instrumentation build only: cycles of the sweeps
#endif
  // estimated cycles of this environment's own solve (the class it files itself under for the next step): fixed build
  // cost + affine arm iterations + serial sweeps x (motor block + generic rows)
  if (lane == 0)
    sm.cost[0] = 20000.f + 150.f * (float)(iters_arm > 0 ? iters_arm : 0) +
                 (float)iters * ((arm_sweep ? 400.f : 0.f) + 55.f * (float)RG);
#ifdef PROFILE_WARM   // probe build: active lanes at entry / after rows+W / after the Delassus block | after look-up / apply / the sweeps
  am5 = __popc(__activemask());
  if (lane == 0) { sm.cost[1] = (float)(am0 + 100 * am1 + 10000 * am2); sm.cost[2] = (float)(am3 + 100 * am4 + 10000 * am5); sm.cost[3] = 0.f; }
#endif
  if (iters_arm > iters) iters = iters_arm;
  sm.mlam[lane] = lane < nd ? m.lam : 0.f;
#pragma unroll
  for (int s = 0; s < NSG; s++) {
    const int gi = GL * s + lane;
    if (gi < RG) sm.glam[gi] = rr.lam[s];
  }
  gsync(g);
  return iters;
}


// ------------------------------------------------------------------------------------------
// Collision stage.  Broadphase: a static candidate-pair list (object x static world, robot proxies x object, robot
// proxies x static world, robot self pairs) culled by bounding spheres.  Narrowphase: closed forms where they are exact
// (cube vertices on the table top while the cube is wholly over it, sphere-box, sphere-plane, sphere-sphere), box-box
// SAT + face clipping (include/b2env_narrowphase.h) for the rest: the cube at the table rim or against a leg, the
// finger-pad boxes against the cube and the table.  Lanes = candidates of one family (cube vertices, spheres, box
// vertices, box pairs, self pairs); every family is compacted into the contact list by ballots, in the canonical order
// of oracle/b2oracle.c collide().  The box-box routine is scalar code run by the lane that owns the pair (rare).
#define B2N_REAL float
#define B2N_FN __device__ __noinline__
#define B2N_SQRT(x) sqrtf(x)
#define B2N_FABS(x) fabsf(x)
#include "../../include/b2env_narrowphase.h"

struct ContactOut {   // what a lane hands to emit_contact
  float pA[3], pB[3], n[3], dist, mu, erp, cfm;
  int key, type, link, link2;
};
__device__ __forceinline__ void emit_contact(EnvSmem& sm, int slot, int maxc, const ContactOut& o) {
  if (slot >= maxc) return;
  Contact& c = sm.con[slot];
  c.key = o.key; c.type = o.type; c.link = o.link; c.link2 = o.link2;
#pragma unroll
  for (int j = 0; j < 3; j++) { c.pA[j] = o.pA[j]; c.pB[j] = o.pB[j]; c.n[j] = o.n[j]; }
  c.dist = o.dist; c.mu = o.mu; c.erp = o.erp;
  sm.con_cfm[slot] = o.cfm;
}
// exclusive prefix over the lanes of the group of a small per-lane count (0..7), and the group total
__device__ __forceinline__ int gprefix3(const Grp& g, int cnt, int& total) {
  const unsigned lt = (1u << g.lane) - 1u;
  const unsigned b0 = gballot(g, cnt & 1), b1 = gballot(g, cnt & 2), b2 = gballot(g, cnt & 4);
  total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
  return __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
}
__device__ __forceinline__ void sbox_get(const b2e_params& P, int k, float* c, float* h) {
  if (P.n_sboxes > 0) {
#pragma unroll
    for (int j = 0; j < 3; j++) { c[j] = P.sbox_c[k][j]; h[j] = P.sbox_h[k][j]; }
  } else {
#pragma unroll
    for (int j = 0; j < 3; j++) { c[j] = 0.5f * (P.table_min[j] + P.table_max[j]); h[j] = 0.5f * (P.table_max[j] - P.table_min[j]); }
  }
}
// sphere vs axis-aligned box (same decisions as oracle sphere_aabox)
__device__ __forceinline__ bool sphere_aabox(const float* c, float r, const float* bc, const float* bh, float margin, float* n,
                                             float* pB, float& dist, bool& top) {
  float l[3], cl[3];
  bool inside = true;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    l[j] = c[j] - bc[j];
    cl[j] = l[j] < -bh[j] ? -bh[j] : (l[j] > bh[j] ? bh[j] : l[j]);
    if (cl[j] != l[j]) inside = false;
  }
  top = false;
  if (!inside) {
    const float dv[3] = {l[0] - cl[0], l[1] - cl[1], l[2] - cl[2]};
    if (dv[0] == 0.f && dv[1] == 0.f && dv[2] > 0.f) {   // over the top face: the closed form of round 1, bit for bit
      dist = c[2] - r - (bc[2] + bh[2]);
      if (!(dist < margin)) return false;
      n[0] = 0.f; n[1] = 0.f; n[2] = 1.f;
      pB[0] = c[0]; pB[1] = c[1]; pB[2] = bc[2] + bh[2];
      top = true;
      return true;
    }
    const float d = sqrtf(dot3(dv, dv));
    dist = d - r;
    if (!(dist < margin)) return false;
#pragma unroll
    for (int j = 0; j < 3; j++) { n[j] = dv[j] / d; pB[j] = bc[j] + cl[j]; }
    return true;
  }
  int ax = 0;
  float best = bh[0] - fabsf(l[0]);
#pragma unroll
  for (int j = 1; j < 3; j++) {
    const float pen = bh[j] - fabsf(l[j]);
    if (pen < best) { best = pen; ax = j; }
  }
  n[0] = n[1] = n[2] = 0.f;
  const float sg = ((ax == 0 ? l[0] : (ax == 1 ? l[1] : l[2])) >= 0.f) ? 1.f : -1.f;
#pragma unroll
  for (int j = 0; j < 3; j++) if (j == ax) { n[j] = sg; cl[j] = sg * bh[j]; }
  dist = -best - r;
#pragma unroll
  for (int j = 0; j < 3; j++) pB[j] = bc[j] + cl[j];
  return true;
}

// capsule shape from its scratch record: p0 (3) | radius | p1 (3) | friction  (same arithmetic as oracle collide())
__device__ __forceinline__ void cap_shape(const float* cs, b2n_shape& cap) {
  const float ax[3] = {cs[4] - cs[0], cs[5] - cs[1], cs[6] - cs[2]};
  const float len = sqrtf(dot3(ax, ax));
  cap.type = B2N_SEGMENT;
#pragma unroll
  for (int j = 0; j < 9; j++) cap.R[j] = 0.f;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    cap.c[j] = 0.5f * (cs[j] + cs[4 + j]);
    cap.R[3 * j + 2] = len > 0.f ? ax[j] / len : (j == 2 ? 1.f : 0.f);
    cap.h[j] = 0.f;
  }
  cap.h[2] = 0.5f * len;
  cap.r = cs[3];
}
// Returns the contact count (<= B2E_MAX_CONTACTS) | overflow << 8.  Reads the link transforms of sm.T; sm.W is scratch here
// (left finite).  Both groups of the warp run every collective together.
__device__ __noinline__ int collide_stage(EnvSmem& sm, const DevModel* __restrict__ M, const DevModelU& U, const b2e_params& P,
                                          unsigned hm, int sh, int lane, float cpx, float cpy, float cpz, float cqx, float cqy,
                                          float cqz, float cqw) {
  const Grp g = {hm, sh, lane};
  const unsigned lt = (1u << lane) - 1u;
  const int maxc = (P.max_contacts > 0 && P.max_contacts < B2E_MAX_CONTACTS) ? P.max_contacts : B2E_MAX_CONTACTS;   // b2e_params.max_contacts
  const float cpos[3] = {cpx, cpy, cpz}, cquat[4] = {cqx, cqy, cqz, cqw};
  float Rc[9];
  quat_to_mat(cquat, Rc);
  const float ca = P.cube_half, margin = P.contact_margin, top = P.table_max[2];
  const float chh[3] = {ca, ca, ca};
  const float rb = ca * 1.7320508075688772f;
  const int nsb = P.n_sboxes > 0 ? P.n_sboxes : 1;
  const float ident[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
  int base = 0, total;
  ContactOut o;
  // ---- 1. cube vs static world ----
  const bool cube_fast = cpos[0] - rb >= P.table_min[0] && cpos[0] + rb <= P.table_max[0] && cpos[1] - rb >= P.table_min[1] &&
                         cpos[1] + rb <= P.table_max[1] && cpos[2] >= top;
  float v_pos[3] = {0.f, 0.f, 0.f};
  if (lane < 8) {
    const float l[3] = {(lane & 1) ? ca : -ca, (lane & 2) ? ca : -ca, (lane & 4) ? ca : -ca};
    m3vec(Rc, l, v_pos);
    v_pos[0] += cpos[0]; v_pos[1] += cpos[1]; v_pos[2] += cpos[2];
  }
  {
    const bool hit = cube_fast && lane < 8 && (v_pos[2] - top) < margin;
    const unsigned b = gballot(g, hit);
    if (hit) {
      o.key = B2E_KEY_CUBE_TABLE + lane; o.type = CT_CUBE_STATIC; o.link = -1; o.link2 = -1;
      o.pA[0] = v_pos[0]; o.pA[1] = v_pos[1]; o.pA[2] = v_pos[2];
      o.pB[0] = v_pos[0]; o.pB[1] = v_pos[1]; o.pB[2] = top;
      o.n[0] = 0.f; o.n[1] = 0.f; o.n[2] = 1.f;
      o.dist = v_pos[2] - top; o.mu = P.cube_mu * P.table_mu; o.erp = P.erp; o.cfm = 0.f;
      emit_contact(sm, base + __popc(b & lt), maxc, o);
    }
    base += __popc(b);
  }
  {  // rim of the top slab, legs: lane = static box, general box-box behind a bounding-sphere cull
    int cnt = 0;
    float nrm[3] = {0.f, 0.f, 1.f};
    b2n_contact pts[B2N_MAX_POINTS];
    const int k = lane;
    if (k < nsb && k >= (cube_fast ? 1 : 0)) {
      float bc[3], bh[3];
      sbox_get(P, k, bc, bh);
      const float d[3] = {cpos[0] - bc[0], cpos[1] - bc[1], cpos[2] - bc[2]};
      if (!(sqrtf(dot3(d, d)) - (rb + sqrtf(dot3(bh, bh))) >= margin)) cnt = b2n_box_box(cpos, Rc, chh, bc, ident, bh, margin, nrm, pts);
    }
    __syncwarp();
    const int off = gprefix3(g, cnt, total);
    for (int q = 0; q < cnt; q++) {
      o.key = B2E_KEY_CUBE_SBOX + B2N_ID_STRIDE * k + pts[q].id; o.type = CT_CUBE_STATIC; o.link = -1; o.link2 = -1;
#pragma unroll
      for (int j = 0; j < 3; j++) { o.pA[j] = pts[q].pa[j]; o.pB[j] = pts[q].pb[j]; o.n[j] = nrm[j]; }
      o.dist = pts[q].dist; o.mu = P.cube_mu * (P.n_sboxes > 0 ? P.sbox_mu[k] : P.table_mu); o.erp = P.erp; o.cfm = 0.f;
      emit_contact(sm, base + off + q, maxc, o);
    }
    base += total;
  }
  {
    const bool hit = !cube_fast && lane < 8 && v_pos[2] < margin;
    const unsigned b = gballot(g, hit);
    if (hit) {
      o.key = B2E_KEY_CUBE_PLANE + lane; o.type = CT_CUBE_STATIC; o.link = -1; o.link2 = -1;
      o.pA[0] = v_pos[0]; o.pA[1] = v_pos[1]; o.pA[2] = v_pos[2];
      o.pB[0] = v_pos[0]; o.pB[1] = v_pos[1]; o.pB[2] = 0.f;
      o.n[0] = 0.f; o.n[1] = 0.f; o.n[2] = 1.f;
      o.dist = v_pos[2]; o.mu = P.cube_mu * P.plane_mu; o.erp = P.erp; o.cfm = 0.f;
      emit_contact(sm, base + __popc(b & lt), maxc, o);
    }
    base += __popc(b);
  }
  // ---- robot proxies in world coordinates: sphere centres (lane = sphere, also staged in sm.W for the self pairs) ----
  const int ns = U.n_spheres;
  float s_c[3] = {0.f, 0.f, 0.f}, s_r = 0.f, s_mu = 0.f, s_erp = P.erp, s_cfm = 0.f;
  int s_link = 0;
  float* scr = sm.W;
  bool s_world = false;   // this lane's sphere is tested against the object and the static world
  if (lane < ns) {
    s_world = !(__ldg(&M->sph_flags[lane]) & 1);
    s_link = __ldg(&M->sph_link[lane]);
    s_r = __ldg(&M->sph_r[lane]);
    s_mu = __ldg(&M->sph_mu[lane]);
    const float se = __ldg(&M->sph_erp[lane]);
    s_erp = se >= 0.f ? se : P.erp;
    s_cfm = __ldg(&M->sph_cfm[lane]);
    const float lc[3] = {__ldg(&M->sph_c[lane][0]), __ldg(&M->sph_c[lane][1]), __ldg(&M->sph_c[lane][2])};
    float Rl[9], oo[3];
#pragma unroll
    for (int k = 0; k < 9; k++) Rl[k] = sm.T[s_link][k];
    m3vec(Rl, lc, oo);
    s_c[0] = sm.T[s_link][9] + oo[0]; s_c[1] = sm.T[s_link][10] + oo[1]; s_c[2] = sm.T[s_link][11] + oo[2];
  }
  scr[lane * 4 + 0] = s_c[0]; scr[lane * 4 + 1] = s_c[1]; scr[lane * 4 + 2] = s_c[2]; scr[lane * 4 + 3] = s_r;
  // finger-pad boxes: world pose (lanes 0..n_boxes-1 own a box; every lane can read them from scratch)
  const int nbx = U.n_boxes;
  float* bscr = scr + 64;     // per box: centre (3) | bounding radius | R (9) | half extents (3)
  if (lane < nbx) {
    const int li = __ldg(&M->box_link[lane]);
    const float lc[3] = {__ldg(&M->box_c[lane][0]), __ldg(&M->box_c[lane][1]), __ldg(&M->box_c[lane][2])};
    float Rl[9], oo[3];
#pragma unroll
    for (int k = 0; k < 9; k++) Rl[k] = sm.T[li][k];
    m3vec(Rl, lc, oo);
    const float h[3] = {__ldg(&M->box_h[lane][0]), __ldg(&M->box_h[lane][1]), __ldg(&M->box_h[lane][2])};
    float* bs = bscr + 16 * lane;
    bs[0] = sm.T[li][9] + oo[0]; bs[1] = sm.T[li][10] + oo[1]; bs[2] = sm.T[li][11] + oo[2];
    bs[3] = sqrtf(dot3(h, h));
#pragma unroll
    for (int k = 0; k < 9; k++) bs[4 + k] = Rl[k];
    bs[13] = h[0]; bs[14] = h[1]; bs[15] = h[2];
  }
  gsync(g);
  // capsules: world end points (lanes 0..n_caps-1 own one; every lane can read them from scratch), then capsule vs cube:
  // lane = capsule, GJK / EPA on the segment core behind a bounding-sphere cull
  const int ncap = U.n_caps;
  float* cscr = bscr + 16 * B2E_MAX_BOXES;   // per capsule: p0 (3) | radius | p1 (3) | friction
  if (lane < ncap) {
    const int li = __ldg(&M->cap_link[lane]);
    float Rl[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Rl[k] = sm.T[li][k];
    float* cs = cscr + 8 * lane;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const float lc[3] = {e ? __ldg(&M->cap_p1[lane][0]) : __ldg(&M->cap_p0[lane][0]), e ? __ldg(&M->cap_p1[lane][1]) : __ldg(&M->cap_p0[lane][1]),
                           e ? __ldg(&M->cap_p1[lane][2]) : __ldg(&M->cap_p0[lane][2])};
      float oo[3];
      m3vec(Rl, lc, oo);
      cs[4 * e + 0] = sm.T[li][9] + oo[0]; cs[4 * e + 1] = sm.T[li][10] + oo[1]; cs[4 * e + 2] = sm.T[li][11] + oo[2];
    }
    cs[3] = __ldg(&M->cap_r[lane]); cs[7] = __ldg(&M->cap_mu[lane]);
  }
  gsync(g);
  // ---- conservative votes of the warp (both environments): does ANY robot proxy come within the margin of the bounding
  //      sphere of the cube / of the top of the static world (every static box and the ground plane lie at z <= ztop)?  If
  //      not, no test of section 2 / section 3 can report a contact (sphere: dist >= gap; box-box: a face axis separates by
  //      more than the margin; GJK: distance >= gap), and the section is skipped — the common case of a rollout ----
  float ztop = 0.f;
  for (int k = 0; k < nsb; k++) {
    float bc[3], bh[3];
    sbox_get(P, k, bc, bh);
    ztop = fmaxf(ztop, bc[2] + bh[2]);
  }
  bool near_cube = false, near_static = false;
  {
    const float rc = rb + margin;
    if (s_world) {
      const float d[3] = {s_c[0] - cpos[0], s_c[1] - cpos[1], s_c[2] - cpos[2]};
      near_cube = near_cube || !(dot3(d, d) > (rc + s_r) * (rc + s_r));
      near_static = near_static || !(s_c[2] - s_r - ztop > margin);
    }
    if (lane < nbx) {
      const float* bs = bscr + 16 * lane;
      const float d[3] = {bs[0] - cpos[0], bs[1] - cpos[1], bs[2] - cpos[2]};
      near_cube = near_cube || !(dot3(d, d) > (rc + bs[3]) * (rc + bs[3]));
      near_static = near_static || !(bs[2] - bs[3] - ztop > margin);
    }
    if (lane < ncap) {
      const float* cs = cscr + 8 * lane;
      const float mid[3] = {0.5f * (cs[0] + cs[4]) - cpos[0], 0.5f * (cs[1] + cs[5]) - cpos[1], 0.5f * (cs[2] + cs[6]) - cpos[2]};
      const float ax[3] = {cs[4] - cs[0], cs[5] - cs[1], cs[6] - cs[2]};
      const float rr = rc + cs[3] + 0.5f * sqrtf(dot3(ax, ax)) * 1.0001f;
      near_cube = near_cube || !(dot3(mid, mid) > rr * rr);
      near_static = near_static || !(fminf(cs[2], cs[6]) - cs[3] - ztop > margin);
    }
  }
  const bool any_cube = __any_sync(FULL, near_cube), any_static = __any_sync(FULL, near_static);
  if (any_cube) {
  // ---- 2. robot vs cube: spheres (closest point on the box) ----
  {
    bool hit = false;
    if (s_world) {
      const float rel[3] = {s_c[0] - cpos[0], s_c[1] - cpos[1], s_c[2] - cpos[2]};
      float l[3], cl[3], nloc[3];
      m3tvec(Rc, rel, l);
      bool inside = true;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        cl[j] = l[j] < -ca ? -ca : (l[j] > ca ? ca : l[j]);
        if (cl[j] != l[j]) inside = false;
      }
      if (!inside) {
        const float dv[3] = {l[0] - cl[0], l[1] - cl[1], l[2] - cl[2]};
        const float d = sqrtf(dot3(dv, dv));
        o.dist = d - s_r;
        hit = o.dist < margin;
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
        o.dist = -best - s_r;
        hit = true;
      }
      if (hit) {
        float pwl[3];
        m3vec(Rc, nloc, o.n);
        m3vec(Rc, cl, pwl);
#pragma unroll
        for (int j = 0; j < 3; j++) { o.pB[j] = cpos[j] + pwl[j]; o.pA[j] = s_c[j] - o.n[j] * s_r; }
      }
    }
    const unsigned b = gballot(g, hit);
    if (hit) {
      o.key = B2E_KEY_SPHERE_CUBE + lane; o.type = CT_ARM_CUBE; o.link = s_link; o.link2 = -1;
      o.mu = P.cube_mu * s_mu; o.erp = s_erp; o.cfm = s_cfm;
      emit_contact(sm, base + __popc(b & lt), maxc, o);
    }
    base += __popc(b);
  }
  if (__any_sync(FULL, nbx > 0)) {  // robot boxes vs cube: lane = box, box-box behind a bounding-sphere cull
    int cnt = 0;
    float nrm[3] = {0.f, 0.f, 1.f};
    b2n_contact pts[B2N_MAX_POINTS];
    if (lane < nbx) {
      const float* bs = bscr + 16 * lane;
      const float bc[3] = {bs[0], bs[1], bs[2]}, bh[3] = {bs[13], bs[14], bs[15]};
      float Rl[9];
#pragma unroll
      for (int k = 0; k < 9; k++) Rl[k] = bs[4 + k];
      const float d[3] = {bc[0] - cpos[0], bc[1] - cpos[1], bc[2] - cpos[2]};
      if (!(sqrtf(dot3(d, d)) - (rb + bs[3]) >= margin)) cnt = b2n_box_box(bc, Rl, bh, cpos, Rc, chh, margin, nrm, pts);
    }
    __syncwarp();
    const int off = gprefix3(g, cnt, total);
    for (int q = 0; q < cnt; q++) {
      const float be = __ldg(&M->box_erp[lane]);
      o.key = B2E_KEY_BOX_CUBE + B2N_ID_STRIDE * lane + pts[q].id; o.type = CT_ARM_CUBE; o.link = __ldg(&M->box_link[lane]); o.link2 = -1;
#pragma unroll
      for (int j = 0; j < 3; j++) { o.pA[j] = pts[q].pa[j]; o.pB[j] = pts[q].pb[j]; o.n[j] = nrm[j]; }
      o.dist = pts[q].dist; o.mu = P.cube_mu * __ldg(&M->box_mu[lane]); o.erp = be >= 0.f ? be : P.erp; o.cfm = __ldg(&M->box_cfm[lane]);
      emit_contact(sm, base + off + q, maxc, o);
    }
    base += total;
  }
  if (__any_sync(FULL, ncap > 0)) {
    bool hit = false;
    if (lane < ncap) {
      b2n_shape cap, cube;
      cap_shape(cscr + 8 * lane, cap);
      const float d[3] = {cap.c[0] - cpos[0], cap.c[1] - cpos[1], cap.c[2] - cpos[2]};
      if (!(sqrtf(dot3(d, d)) - (rb + cap.h[2] + cap.r) >= margin)) {
        cube.type = B2N_BOX; cube.r = 0.f;
#pragma unroll
        for (int j = 0; j < 3; j++) { cube.c[j] = cpos[j]; cube.h[j] = ca; }
#pragma unroll
        for (int j = 0; j < 9; j++) cube.R[j] = Rc[j];
        b2n_contact pt;
        float nn[3];
        if (b2n_convex_contact(&cap, &cube, margin, nn, &pt)) {
          hit = true;
#pragma unroll
          for (int j = 0; j < 3; j++) { o.pA[j] = pt.pa[j]; o.pB[j] = pt.pb[j]; o.n[j] = nn[j]; }
          o.dist = pt.dist;
        }
      }
    }
    const unsigned b = gballot(g, hit);
    if (hit) {
      o.key = B2E_KEY_CAP_CUBE + lane; o.type = CT_ARM_CUBE; o.link = __ldg(&M->cap_link[lane]); o.link2 = -1;
      o.mu = P.cube_mu * cscr[8 * lane + 7]; o.erp = P.erp; o.cfm = 0.f;
      emit_contact(sm, base + __popc(b & lt), maxc, o);
    }
    base += __popc(b);
  }
  }
  if (any_static) {
  // ---- 3. robot vs static world: spheres vs static boxes (lane = sphere, at most three boxes each), vs the ground plane ----
  {
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
    const int off = gprefix3(g, cnt, total);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (c < cnt) {
        const int k = kk[c];
        o.key = (k == 0 && tp[c]) ? B2E_KEY_SPHERE_TABLE + lane : B2E_KEY_SPHERE_SBOX + 8 * lane + k;
        o.type = CT_ARM_STATIC; o.link = s_link; o.link2 = -1;
#pragma unroll
        for (int j = 0; j < 3; j++) { o.n[j] = nn[c][j]; o.pB[j] = pb[c][j]; o.pA[j] = s_c[j] - nn[c][j] * s_r; }
        o.dist = dd[c]; o.mu = (P.n_sboxes > 0 ? P.sbox_mu[k] : P.table_mu) * s_mu; o.erp = s_erp; o.cfm = s_cfm;
        emit_contact(sm, base + off + c, maxc, o);
      }
    }
    base += total;
  }
  {
    const float dist = s_c[2] - s_r;
    const bool hit = s_world && dist < margin;
    const unsigned b = gballot(g, hit);
    if (hit) {
      o.key = B2E_KEY_SPHERE_PLANE + lane; o.type = CT_ARM_STATIC; o.link = s_link; o.link2 = -1;
      o.n[0] = 0.f; o.n[1] = 0.f; o.n[2] = 1.f;
      o.pA[0] = s_c[0]; o.pA[1] = s_c[1]; o.pA[2] = dist;
      o.pB[0] = s_c[0]; o.pB[1] = s_c[1]; o.pB[2] = 0.f;
      o.dist = dist; o.mu = P.plane_mu * s_mu; o.erp = s_erp; o.cfm = s_cfm;
      emit_contact(sm, base + __popc(b & lt), maxc, o);
    }
    base += __popc(b);
  }
  if (__any_sync(FULL, nbx > 0)) {
    // finger-pad boxes: lane = 8 * box + vertex for the vertex families, lane = 8 * box + static box for the general pairs
    const int bxi = lane >> 3, sub = lane & 7;
    const bool has = bxi < nbx;
    const float* bs = bscr + 16 * (has ? bxi : 0);
    const float bc[3] = {bs[0], bs[1], bs[2]}, bh[3] = {bs[13], bs[14], bs[15]}, br = bs[3];
    float Rl[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Rl[k] = bs[4 + k];
    const bool bfast = bc[0] - br >= P.table_min[0] && bc[0] + br <= P.table_max[0] && bc[1] - br >= P.table_min[1] &&
                       bc[1] + br <= P.table_max[1] && bc[2] >= top;
    float bv[3];
    {
      const float l[3] = {(sub & 1) ? bh[0] : -bh[0], (sub & 2) ? bh[1] : -bh[1], (sub & 4) ? bh[2] : -bh[2]};
      m3vec(Rl, l, bv);
      bv[0] += bc[0]; bv[1] += bc[1]; bv[2] += bc[2];
    }
    const int blink = has ? __ldg(&M->box_link[bxi]) : 0;
    const float bmu = has ? __ldg(&M->box_mu[bxi]) : 0.f, bcfm = has ? __ldg(&M->box_cfm[bxi]) : 0.f;
    const float be = has ? __ldg(&M->box_erp[bxi]) : -1.f, berp = be >= 0.f ? be : P.erp;
    {
      const bool hit = has && bfast && (bv[2] - top) < margin;
      const unsigned b = gballot(g, hit);
      if (hit) {
        o.key = B2E_KEY_BOXV_TABLE + lane; o.type = CT_ARM_STATIC; o.link = blink; o.link2 = -1;
        o.pA[0] = bv[0]; o.pA[1] = bv[1]; o.pA[2] = bv[2];
        o.pB[0] = bv[0]; o.pB[1] = bv[1]; o.pB[2] = top;
        o.n[0] = 0.f; o.n[1] = 0.f; o.n[2] = 1.f;
        o.dist = bv[2] - top; o.mu = P.table_mu * bmu; o.erp = berp; o.cfm = bcfm;
        emit_contact(sm, base + __popc(b & lt), maxc, o);
      }
      base += __popc(b);
    }
    {
      int cnt = 0;
      float nrm[3] = {0.f, 0.f, 1.f};
      b2n_contact pts[B2N_MAX_POINTS];
      const int k = sub;
      if (has && k < nsb && k >= (bfast ? 1 : 0)) {
        float sc2[3], sh2[3];
        sbox_get(P, k, sc2, sh2);
        const float d[3] = {bc[0] - sc2[0], bc[1] - sc2[1], bc[2] - sc2[2]};
        if (!(sqrtf(dot3(d, d)) - (br + sqrtf(dot3(sh2, sh2))) >= margin)) cnt = b2n_box_box(bc, Rl, bh, sc2, ident, sh2, margin, nrm, pts);
      }
      __syncwarp();
      const int off = gprefix3(g, cnt, total);
      for (int q = 0; q < cnt; q++) {
        o.key = B2E_KEY_BOX_SBOX + B2N_ID_STRIDE * lane + pts[q].id; o.type = CT_ARM_STATIC; o.link = blink; o.link2 = -1;
#pragma unroll
        for (int j = 0; j < 3; j++) { o.pA[j] = pts[q].pa[j]; o.pB[j] = pts[q].pb[j]; o.n[j] = nrm[j]; }
        o.dist = pts[q].dist; o.mu = (P.n_sboxes > 0 ? P.sbox_mu[k] : P.table_mu) * bmu; o.erp = berp; o.cfm = bcfm;
        emit_contact(sm, base + off + q, maxc, o);
      }
      base += total;
    }
    {
      const bool hit = has && !bfast && bv[2] < margin;
      const unsigned b = gballot(g, hit);
      if (hit) {
        o.key = B2E_KEY_BOXV_PLANE + lane; o.type = CT_ARM_STATIC; o.link = blink; o.link2 = -1;
        o.pA[0] = bv[0]; o.pA[1] = bv[1]; o.pA[2] = bv[2];
        o.pB[0] = bv[0]; o.pB[1] = bv[1]; o.pB[2] = 0.f;
        o.n[0] = 0.f; o.n[1] = 0.f; o.n[2] = 1.f;
        o.dist = bv[2]; o.mu = P.plane_mu * bmu; o.erp = berp; o.cfm = bcfm;
        emit_contact(sm, base + __popc(b & lt), maxc, o);
      }
      base += __popc(b);
    }
  }
  if (__any_sync(FULL, ncap > 0)) {
    // capsules vs static boxes: lane = 8 * capsule + static box (two capsules per pass of 16 lanes), GJK / EPA behind an
    // axis-aligned bounding-box cull; then vs the ground plane: lane = 2 * capsule + end (closed form)
    for (int c0 = 0; c0 < ncap; c0 += 2) {   // warp-uniform trip count: both groups share the model
      const int c = c0 + (lane >> 3), k = lane & 7;
      bool hit = false;
      if (c < ncap && k < nsb) {
        const float* cs = cscr + 8 * c;
        float bc[3], bh[3];
        sbox_get(P, k, bc, bh);
        bool apart = false;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const float lo = fminf(cs[j], cs[4 + j]), hi = fmaxf(cs[j], cs[4 + j]);
          if (lo - cs[3] - (bc[j] + bh[j]) >= margin || (bc[j] - bh[j]) - (hi + cs[3]) >= margin) apart = true;
        }
        if (!apart) {
          b2n_shape cap, box;
          cap_shape(cs, cap);
          box.type = B2N_BOX; box.r = 0.f;
#pragma unroll
          for (int j = 0; j < 3; j++) { box.c[j] = bc[j]; box.h[j] = bh[j]; }
#pragma unroll
          for (int j = 0; j < 9; j++) box.R[j] = ident[j];
          b2n_contact pt;
          float nn[3];
          if (b2n_convex_contact(&cap, &box, margin, nn, &pt)) {
            hit = true;
#pragma unroll
            for (int j = 0; j < 3; j++) { o.pA[j] = pt.pa[j]; o.pB[j] = pt.pb[j]; o.n[j] = nn[j]; }
            o.dist = pt.dist;
          }
        }
      }
      const unsigned b = gballot(g, hit);
      if (hit) {
        o.key = B2E_KEY_CAP_SBOX + 8 * c + k; o.type = CT_ARM_STATIC; o.link = __ldg(&M->cap_link[c]); o.link2 = -1;
        o.mu = (P.n_sboxes > 0 ? P.sbox_mu[k] : P.table_mu) * cscr[8 * c + 7]; o.erp = P.erp; o.cfm = 0.f;
        emit_contact(sm, base + __popc(b & lt), maxc, o);
      }
      base += __popc(b);
    }
    {
      const int c = lane >> 1, e = lane & 1;
      const bool has = c < ncap;
      const float* cs = cscr + 8 * (has ? c : 0);
      const float dist = cs[4 * e + 2] - cs[3];
      const bool hit = has && dist < margin;
      const unsigned b = gballot(g, hit);
      if (hit) {
        o.key = B2E_KEY_CAP_PLANE + lane; o.type = CT_ARM_STATIC; o.link = __ldg(&M->cap_link[c]); o.link2 = -1;
        o.n[0] = 0.f; o.n[1] = 0.f; o.n[2] = 1.f;
        o.pA[0] = cs[4 * e]; o.pA[1] = cs[4 * e + 1]; o.pA[2] = dist;
        o.pB[0] = cs[4 * e]; o.pB[1] = cs[4 * e + 1]; o.pB[2] = 0.f;
        o.dist = dist; o.mu = P.plane_mu * cs[7]; o.erp = P.erp; o.cfm = 0.f;
        emit_contact(sm, base + __popc(b & lt), maxc, o);
      }
      base += __popc(b);
    }
  }
  }
  // ---- 4. robot self-collision (URDF_USE_SELF_COLLISION, panda_env.py:53): sphere pairs of non-neighbouring links ----
  {
    const int nsp = U.n_self_pairs;
    for (int p0 = 0; p0 < nsp; p0 += GL) {   // warp-uniform trip count: both groups share the model
      const int p = p0 + lane;
      bool hit = false;
      if (p < nsp) {
        const int sa = __ldg(&M->self_a[p]), sb = __ldg(&M->self_b[p]);
        const float ra = scr[sa * 4 + 3], rbb = scr[sb * 4 + 3];
        const float d[3] = {scr[sa * 4] - scr[sb * 4], scr[sa * 4 + 1] - scr[sb * 4 + 1], scr[sa * 4 + 2] - scr[sb * 4 + 2]};
        const float dn = sqrtf(dot3(d, d));
        o.dist = dn - (ra + rbb);
        hit = o.dist < margin && dn > 1e-9f;
        if (hit) {
          o.key = B2E_KEY_SELF + p; o.type = CT_ARM_ARM; o.link = __ldg(&M->sph_link[sa]); o.link2 = __ldg(&M->sph_link[sb]);
#pragma unroll
          for (int j = 0; j < 3; j++) {
            o.n[j] = d[j] / dn;
            o.pA[j] = scr[sa * 4 + j] - o.n[j] * ra;
            o.pB[j] = scr[sb * 4 + j] + o.n[j] * rbb;
          }
          o.mu = __ldg(&M->sph_mu[sa]) * __ldg(&M->sph_mu[sb]); o.erp = P.erp; o.cfm = 0.f;
        }
      }
      const unsigned b = gballot(g, hit);
      if (hit) emit_contact(sm, base + __popc(b & lt), maxc, o);
      base += __popc(b);
    }
  }
  gsync(g);
  // leave the scratch finite for the solver tables (they are zeroed at kernel start and must stay finite)
  for (int k = lane; k < 64 + 16 * B2E_MAX_BOXES + 8 * B2E_MAX_CAPS; k += GL) scr[k] = 0.f;
  gsync(g);
  return (base > maxc ? maxc : base) | (base > maxc ? 256 : 0);
}

// ------------------------------------------------------------------------------------------
// the fused step kernel
template <bool IK>
__global__ void __launch_bounds__(32 * WPB, MINB)
step_kernel(const DevModel* __restrict__ M, const __grid_constant__ DevModelU U, const __grid_constant__ b2e_params P, DevState st, const float* __restrict__ action,
            float* __restrict__ obs_out, float* __restrict__ reward_out, float* __restrict__ done_out, int nsub,
            int mode, int record_contacts, const int* __restrict__ env_ids, int n_ids, int env_offset, int* sched, int seq,
            int role, int nslot, int tail_cap, int tail_cost) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wpb = blockDim.x >> 5;                          // warps per block: WPB (main / plain launches) or the tail's
  const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
  const int half = wl >> 4, lane = wl & (GL - 1);          // two environments per warp, 16 lanes each
  const Grp g = {0xffffu << (GL * half), GL * half, lane};
  // slot -> environment: the whole batch in cost order (ROLE_MAIN / ROLE_TAIL, see NBK_MAIN), or identity / the subset
  // listed in env_ids (ROLE_PLAIN: per-env resets (row f1), chunked host path)
  int n_slots = n_ids;   // env_ids: listed envs; else the contiguous range [env_offset, env_offset + n_ids)
  // main / plain launches: consecutive slots per block.  Tail launch: slot s -> block s % grid, warp (s / grid) % wpb, group
  // s / (grid * wpb), so that a short tail list is spread one environment per block (each block has an SM to itself, see
  // launch_step) before a second warp of any block is used, and the second group of a warp last of all
  const int slot = role == ROLE_TAIL ? (half * wpb + warp) * (int)gridDim.x + (int)blockIdx.x : (blockIdx.x * wpb + warp) * 2 + half;
  int env;
  if (role == ROLE_PLAIN) {
    const int sc = slot < n_slots ? slot : n_slots - 1;
    env = env_ids ? env_ids[sc] : env_offset + sc;
  } else {
    const int qp = (seq - 1) & 3, pp = (seq - 1) & 1, B = st.B;
    if (role == ROLE_MAIN && blockIdx.x == 0 && (int)threadIdx.x < NBK)
      sched[SCHED_CNT((seq + 1) & 3, threadIdx.x)] = 0;   // the next step's counters (not used by this step's launches)
    if (role == ROLE_TAIL) {
      n_slots = min(sched[SCHED_CNT(qp, NBK_MAIN)], tail_cap);
      // a warp whose first group has no environment leaves (block barriers count the warps that are left); a second group
      // without one shadows the first (stores masked)
      const int slot_a = warp * (int)gridDim.x + (int)blockIdx.x;
      if (slot_a >= n_slots) {   // (after zeroing this thread's share of the block's overflow slots, see below)
        float4* s4 = reinterpret_cast<float4*>(smem_raw + sizeof(EnvSmem) * 2 * wpb);
        for (int k = threadIdx.x; k < (int)(sizeof(BigSlot) / 16) * nslot; k += blockDim.x) s4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
      }
      env = sched[SCHED_TAIL(pp, B) + (slot < n_slots ? slot : slot_a)];
    } else {
      int cnt[NBK_MAIN];
#pragma unroll
      for (int b = 0; b < NBK_MAIN; b++) cnt[b] = sched[SCHED_CNT(qp, b)];
      n_slots = 0;
#pragma unroll
      for (int b = 0; b < NBK_MAIN; b++) n_slots += cnt[b];
      if ((int)blockIdx.x * 2 * wpb >= n_slots) return;
      int rem = slot < n_slots ? slot : n_slots - 1, cls = 0;   // heaviest class first
#pragma unroll
      for (int b = NBK_MAIN - 1; b >= 0; b--) {
        if (rem >= 0 && rem < cnt[b]) { cls = b; break; }
        rem -= cnt[b];
      }
      env = sched[SCHED_LIST(pp, cls, B) + rem];
    }
  }
  const bool live_env = slot < n_slots;      // padding warps of the last block shadow the last slot, stores masked
  EnvSmem& sm = reinterpret_cast<EnvSmem*>(smem_raw)[warp * 2 + half];
  const int slots_off = (int)sizeof(EnvSmem) * 2 * wpb;
  BigSlot* slots = reinterpret_cast<BigSlot*>(smem_raw + slots_off);
  int* slot_owner = reinterpret_cast<int*>(smem_raw + slots_off + sizeof(BigSlot) * nslot);   // [16] owners | [NBK] class
  int* cls_hist = slot_owner + 16;                                                              // histogram | [NBK] bases
  int* cls_base = cls_hist + NBK;
  if ((int)threadIdx.x < NBK) cls_hist[threadIdx.x] = 0;   // ordered before its use by the barriers of the step loop
  const int nd = U.n_dof, nl = U.n_links;
  const float dt = P.dt;

  // ---- load state (env-major field groups; lanes = dofs for q/qd/targets) ----
  const bool is_dof = lane < nd;
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
  float my_act = 0.f;
  if (mode == B2E_MODE_ACTION && lane < P.n_act) my_act = action[env * P.n_act + lane];
  float my_hp = (IK && lane < 6) ? st.hand_pose[env * 6 + lane] : 0.f;   // commanded hand pose (lane = component)
  const float my_lower = is_dof ? __ldg(&M->lower[lane]) : 0.f, my_upper = is_dof ? __ldg(&M->upper[lane]) : 0.f;
  const bool ctrl_gains = !IK && mode != B2E_MODE_HOLD && mode != B2E_MODE_IK_POSE;
  // Cartesian control with a velocity cap (robot.apply_action(action, max_vel), panda_env.py:285-291): the controlled joints
  // run with PyBullet's default position gain and maxVelocity = max_vel
  const bool ik_cap = IK && (mode == B2E_MODE_ACTION || mode == B2E_MODE_IK_POSE) && P.ik_max_vel > 0.f && lane < P.n_ctrl;
  const float my_kp = ik_cap ? P.kp_ik_max_vel
                      : ((ctrl_gains && lane < P.n_ctrl) ? P.kp_ctrl
                         : ((ctrl_gains && P.task == B2E_TASK_GRASP && lane >= P.n_ctrl) ? P.kp_grip : P.kp_hold));
  const float my_mv = ik_cap ? P.ik_max_vel : (is_dof ? __ldg(&M->max_vel[lane]) : -1.f);
  const float grip_cmd = SHF(my_act, P.n_ctrl & (GL - 1));   // GRASP: action[n_ctrl] = gripper command
  int iters = 0, nc = 0, R = 0;
#ifdef PROFILE_CYCLES
  int prof_pgs = 0;
#endif
  bool stop = false;
#ifdef PROFILE_STAGES   // instrumentation build only: cycle stamps of the stages of the last sub-step -> B2E_F_CONTACTS[env][0..7]
  long long prof_t[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int prof_slot = -2, prof_mate = -1, prof_glob = 0, prof_rgw = 0;
#ifdef PROFILE_WARM   // + active lanes at every stamp -> B2E_F_CONTACTS[env][32..45]
  int prof_am[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PROF_T(k) do { prof_t[k] = clock64(); prof_am[k] = __popc(__activemask()); } while (0)
#else
#define PROF_T(k) prof_t[k] = clock64()
#endif
#else
#define PROF_T(k) do { } while (0)
#endif
  PROF_T(0);
  {
    // solver tables start finite (see motor_step / generic_step): A | W | WT of this environment and the overflow slots
    float4* t4 = reinterpret_cast<float4*>(&sm);
    for (int k = lane; k < (int)(offsetof(EnvSmem, T) / 16); k += GL) t4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* s4 = reinterpret_cast<float4*>(slots);
    for (int k = threadIdx.x; k < (int)(sizeof(BigSlot) / 16) * nslot; k += blockDim.x) s4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  gsync(g);

  const int li = lane < nl ? lane : 0;
  const int my_dof = __ldg(&M->dof[li]);
  const bool link_has_dof = lane < nl && my_dof >= 0;
  float Rm[9], pw[3];
  for (int sub = 0;; sub++) {
    // ---- forward kinematics (lane = link) of the current q: start-of-step kinematics of this
    //      sub-step, and at the same time the post-step kinematics of the previous one ----
    {
      const float qi = SHF(my_q, my_dof < 0 ? 0 : my_dof);
      fk_lanes(M, U, g, link_has_dof ? qi : 0.f, Rm, pw);
    }
    // ---- termination inside apply_action (panda_push_gym_env.py:239-242), for the previous sub-step ----
    {
      // EE position for the reach test: the shuffles run unconditionally (both groups together)
      float e3[3] = {0.f, 0.f, 0.f};
      if (P.task == B2E_TASK_REACH) {
        const int ee = U.ee_link;
        float cm[3] = {U.ee_com[0], U.ee_com[1], U.ee_com[2]}, o[3];
        m3vec(Rm, cm, o);
        e3[0] = SHF(pw[0] + o[0], ee); e3[1] = SHF(pw[1] + o[1], ee); e3[2] = SHF(pw[2] + o[2], ee);
      }
      if (sub > 0 && mode == B2E_MODE_ACTION && !stop) {
        float d;
        if (P.task == B2E_TASK_PUSH) {
          float dd[3] = {cpos[0] - target[0], cpos[1] - target[1], cpos[2] - target[2]};
          d = sqrtf(dot3(dd, dd));
        } else if (P.task == B2E_TASK_GRASP) {
          d = (cpos[2] - P.grasp_rest_z >= P.grasp_lift) ? 0.f : 2.f * P.dist_min + 1.f;
        } else {
          float dd[3] = {e3[0] - cpos[0], e3[1] - cpos[1], e3[2] - cpos[2]};
          d = sqrtf(dot3(dd, dd));
        }
        if (P.goal_env) { if (counter > P.max_steps) stop = true; else counter++; }
        else if (d <= P.dist_min) { terminated = 1; stop = true; }
        else if (terminated || counter > P.max_steps) stop = true;
        else counter++;
      }
    }
    if (sub >= nsub) break;
    PROF_T(1);   // start of the sub-step (after FK + termination bookkeeping)
    const bool ghost = stop;   // terminated mid-repeat (panda_push_gym_env.py:239-240): keep pace with the block, change nothing
    if ((int)threadIdx.x < nslot) slot_owner[threadIdx.x] = -1;   // overflow slots are free again (claimed after the barriers below)
    __syncthreads();

    // ---- action -> motor targets (panda_push_gym_env.py:225-230, panda_env.py:303) ----
    if (!IK && mode == B2E_MODE_ACTION && lane < P.n_ctrl && !ghost) {
      my_act *= P.act_scale;
      my_target = fminf(fmaxf(my_q + my_act, my_lower), my_upper);
    }
    if (!IK && mode == B2E_MODE_ACTION && P.task == B2E_TASK_GRASP && is_dof && lane >= P.n_ctrl && !ghost)
      my_target = fminf(fmaxf(0.02f + 0.02f * grip_cmd, my_lower), my_upper);
    // ---- Cartesian control (panda_push_gym_env.py:197-222, panda_env.py:229-282): hand pose += scaled
    //      action, clamps, IK -> position targets of every movable joint ----
    if (IK && (mode == B2E_MODE_ACTION || mode == B2E_MODE_IK_POSE)) {
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
        tp[k] = SHF(my_hp, k);
        const float e = P.ik_orientation ? SHF(my_hp, 3 + k) : P.home_hand_pose[3 + k];
        eu[k] = fminf(fmaxf(e, -3.14159265358979323846f), 3.14159265358979323846f);
      }
      tp[2] = fminf(fmaxf(tp[2], P.ws_lim[2][0]), P.ws_lim[2][1]);
      euler_to_quat(eu, tq);
      const float t = ik_solve(sm.W, M, U, P.ik_iters, P.ik_residual, P.ik_damping, g.hm, g.sh, lane, my_q, tp[0], tp[1], tp[2],
                               tq[0], tq[1], tq[2], tq[3]);
      if (is_dof && !ghost) my_target = t;
    }
    const float qdi_raw = SHF(my_qd, my_dof < 0 ? 0 : my_dof);
    const float qdi = link_has_dof ? qdi_raw : 0.f;
    if (lane < nl) {
#pragma unroll
      for (int k = 0; k < 9; k++) sm.T[lane][k] = Rm[k];
#pragma unroll
      for (int k = 0; k < 3; k++) sm.T[lane][9 + k] = pw[k];
    }

    // ---- dynamics in world coordinates about O = base position ----
    // spatial axis S = (w ; v_O), link spatial inertia (m, h = m c, I_O sym6)
    float S[6] = {0, 0, 0, 0, 0, 0};
    float Iw[16];  // f(6) | m | h(3) | I_O(6): the subtree-summed block
    {
      const int jt = __ldg(&M->jtype[li]);
      float ax[3] = {__ldg(&M->axis[li][0]), __ldg(&M->axis[li][1]), __ldg(&M->axis[li][2])}, aw[3];
      m3vec(Rm, ax, aw);
      float rel[3] = {pw[0] - U.base_pos[0], pw[1] - U.base_pos[1], pw[2] - U.base_pos[2]};
      if (lane < nl && jt == B2E_JOINT_REVOLUTE) {
        float t[3];
        cross3(rel, aw, t);
        S[0] = aw[0]; S[1] = aw[1]; S[2] = aw[2]; S[3] = t[0]; S[4] = t[1]; S[5] = t[2];
      } else if (lane < nl && jt == B2E_JOINT_PRISMATIC) {
        S[3] = aw[0]; S[4] = aw[1]; S[5] = aw[2];
      }
      float cm[3] = {__ldg(&M->com[li][0]), __ldg(&M->com[li][1]), __ldg(&M->com[li][2])}, c[3];
      m3vec(Rm, cm, c);
      c[0] += rel[0]; c[1] += rel[1]; c[2] += rel[2];
      const float ms = lane < nl ? __ldg(&M->mass[li]) : 0.f;
      float Ic[9], RI[9], Icw[9];
#pragma unroll
      for (int k = 0; k < 9; k++) Ic[k] = lane < nl ? __ldg(&M->inertia[li][k]) : 0.f;
      m3mul(Rm, Ic, RI);
      // Icw = RI * Rm^T
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) Icw[3 * i + j] = RI[3 * i] * Rm[3 * j] + RI[3 * i + 1] * Rm[3 * j + 1] + RI[3 * i + 2] * Rm[3 * j + 2];
      const float cc = dot3(c, c);
      Iw[6] = ms;
      Iw[7] = ms * c[0]; Iw[8] = ms * c[1]; Iw[9] = ms * c[2];
      Iw[10] = Icw[0] + ms * (cc - c[0] * c[0]);  // xx
      Iw[11] = Icw[1] - ms * c[0] * c[1];         // xy
      Iw[12] = Icw[2] - ms * c[0] * c[2];         // xz
      Iw[13] = Icw[4] + ms * (cc - c[1] * c[1]);  // yy
      Iw[14] = Icw[5] - ms * c[1] * c[2];         // yz
      Iw[15] = Icw[8] + ms * (cc - c[2] * c[2]);  // zz
    }
    // link velocities v = sum over path of S*qd ; bias accelerations
    float vJ[6], v[6];
#pragma unroll
    for (int k = 0; k < 6; k++) { vJ[k] = S[k] * qdi; v[k] = vJ[k]; }
    path_sum6(M, U, g, v);
    float ab[6];
    {
      float a[3], b[3], c2[3];
      cross3(v, vJ, a);
      cross3(v, vJ + 3, b);
      cross3(v + 3, vJ, c2);
      ab[0] = a[0]; ab[1] = a[1]; ab[2] = a[2];
      ab[3] = b[0] + c2[0]; ab[4] = b[1] + c2[1]; ab[5] = b[2] + c2[2];
    }
    path_sum6(M, U, g, ab);
    ab[3] -= P.gravity[0]; ab[4] -= P.gravity[1]; ab[5] -= P.gravity[2];  // a_0 = -g
    {
      // f = I*ab + v x* (I*v)
      const float ms = Iw[6];
      const float* h = &Iw[7];
      const float* I = &Iw[10];
      float Ia_n[3], Ia_f[3], Iv_n[3], Iv_f[3], t[3];
      // I*x: n = I_O w + h x vlin ; f = m vlin - h x w
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
    // child -> parent accumulation of [f | composite inertia]: host-built schedule (heavy-path suffix
    // scans + folds of light subtrees); every round each lane pulls one source lane (15 = a zero lane)
    {
      const unsigned sch0 = __ldg(&M->acc_sched[0][lane]), sch1 = __ldg(&M->acc_sched[1][lane]);
      const int rounds = U.acc_rounds;
      for (int rd = 0; rd < rounds; rd++) {
        const int src = rd < 8 ? ((sch0 >> (4 * rd)) & 15) : ((sch1 >> (4 * (rd - 8))) & 15);
#pragma unroll
        for (int k = 0; k < 16; k++) Iw[k] += SHF(Iw[k], src);
      }
    }
    // move to dof lanes: lane d gets S, f^c, I^c of its link
    const int lk = is_dof ? __ldg(&M->dof_link[lane]) : 0;
    float Sd[6], Fc[6], Ic[10];
#pragma unroll
    for (int k = 0; k < 6; k++) { Sd[k] = SHF(S[k], lk); Fc[k] = SHF(Iw[k], lk); }
#pragma unroll
    for (int k = 0; k < 10; k++) Ic[k] = SHF(Iw[6 + k], lk);
    if (is_dof) {
#pragma unroll
      for (int k = 0; k < 6; k++) sm.S[lane][k] = Sd[k];
    }
    // F = I^c S_d ; tau_bias = S_d . f^c
    float F[6];
    {
      const float ms = Ic[0];
      const float* h = &Ic[1];
      const float* I = &Ic[4];
      float t[3];
      cross3(h, Sd + 3, t);
      F[0] = I[0] * Sd[0] + I[1] * Sd[1] + I[2] * Sd[2] + t[0];
      F[1] = I[1] * Sd[0] + I[3] * Sd[1] + I[4] * Sd[2] + t[1];
      F[2] = I[2] * Sd[0] + I[4] * Sd[1] + I[5] * Sd[2] + t[2];
      cross3(h, Sd, t);
      F[3] = ms * Sd[3] - t[0]; F[4] = ms * Sd[4] - t[1]; F[5] = ms * Sd[5] - t[2];
    }
    const float tau_b = Sd[0] * Fc[0] + Sd[1] * Fc[1] + Sd[2] * Fc[2] + Sd[3] * Fc[3] + Sd[4] * Fc[4] + Sd[5] * Fc[5];
    // joint-space inertia: M[d][e] = S_e . F_d for e on the path to d; symmetrised through smem
    const unsigned my_mask = is_dof ? __ldg(&M->link_dofmask[lk]) : 0u;
    for (int k = lane; k < NDMAX * (NDMAX + 1); k += GL) (&sm.Minv[0][0])[k] = 0.f;
    gsync(g);
#pragma unroll
    for (int e = 0; e < NDMAX; e++) {
      float Se[6];
#pragma unroll
      for (int k = 0; k < 6; k++) Se[k] = SHF(Sd[k], e);
      const float val = Se[0] * F[0] + Se[1] * F[1] + Se[2] * F[2] + Se[3] * F[3] + Se[4] * F[4] + Se[5] * F[5];
      if (is_dof && e < nd && ((my_mask >> e) & 1)) {
        sm.Minv[lane][e] = val;
        sm.Minv[e][lane] = val;
      }
    }
    gsync(g);
    float a[NDMAX];
#pragma unroll
    for (int e = 0; e < NDMAX; e++) a[e] = (is_dof && e < nd) ? sm.Minv[lane][e] : ((e == lane) ? 1.f : 0.f);
    // in-place Gauss-Jordan inverse, lane = row
#pragma unroll
    for (int k = 0; k < NDMAX; k++) {
      float rk[NDMAX];
#pragma unroll
      for (int j = 0; j < NDMAX; j++) rk[j] = SHF(a[j], k);
      const float pinv = 1.0f / rk[k];
      if (lane == k) {
#pragma unroll
        for (int j = 0; j < NDMAX; j++) a[j] = (j == k) ? pinv : rk[j] * pinv;
      } else {
        const float f = a[k];
#pragma unroll
        for (int j = 0; j < NDMAX; j++) a[j] = (j == k) ? -f * pinv : fmaf(-f * pinv, rk[j], a[j]);
      }
    }
    gsync(g);
    if (lane < NDMAX) {
#pragma unroll
      for (int e = 0; e < NDMAX; e++) sm.Minv[lane][e] = (lane < nd && e < nd) ? a[e] : 0.f;
    }
    // unconstrained acceleration and v*
    const float rhs_d = is_dof ? (-tau_b - __ldg(&M->joint_damping[lane]) * my_qd) : 0.f;
    float qdd = 0.f;
#pragma unroll
    for (int e = 0; e < NDMAX; e++) qdd = fmaf(a[e], SHF(rhs_d, e), qdd);
    const float vstar_d = my_qd + dt * qdd;
    float cvs[3], cws[3];
    {
      const float vn = sqrtf(dot3(cv, cv)), wn = sqrtf(dot3(cw, cw));
#pragma unroll
      for (int k = 0; k < 3; k++) {
        cvs[k] = cv[k] + dt * (P.gravity[k] - cv[k] * (P.damp_lin_k1 + P.damp_lin_k2 * vn));
        cws[k] = cw[k] + dt * (-cw[k] * (P.damp_ang_k1 + P.damp_ang_k2 * wn));
      }
    }
    if (is_dof) sm.vstar[lane] = vstar_d;
    if (lane >= NDMAX && lane < 16) {
      const int k = lane - NDMAX;
      sm.vstar[lane] = k < 3 ? cvs[k] : (k < 6 ? cws[k - 3] : 0.f);
    }
    if (lane < NDMAX && !is_dof) sm.vstar[lane] = 0.f;

    PROF_T(2);   // dynamics done
    PHASE_BARRIER();
    // ---- collision detection (pre-step poses): broadphase + narrowphase, see collide_stage ----
    {
      const int cr = collide_stage(sm, M, U, P, g.hm, g.sh, lane, cpos[0], cpos[1], cpos[2], cquat[0], cquat[1], cquat[2], cquat[3]);
      nc = cr & 255;
      if (cr & 256) flags |= B2E_ST_CONTACT_OVERFLOW;
    }
#ifdef PROFILE_WARM
    const int amA = __popc(__activemask());
#endif
    const unsigned lt = (1u << lane) - 1u;
    // joint-limit rows near a limit (lane = dof): order (dof, lower) then (dof, upper)
    const float dlo = my_q - my_lower, dup = my_upper - my_q;
    const float lmar = is_dof ? __ldg(&M->limit_margin[lane]) : 0.f;
    const bool lo_hit = is_dof && dlo < lmar, up_hit = is_dof && dup < lmar;
    const unsigned blo = gballot(g, lo_hit), bup = gballot(g, up_hit);
    int nlim = __popc(blo) + __popc(bup);
    if (nlim > B2E_MAX_LIMROWS) { flags |= B2E_ST_LIMIT_OVERFLOW; nlim = B2E_MAX_LIMROWS; }
    if (lo_hit) {
      const int slot = __popc(blo & lt) + __popc(bup & lt);
      if (slot < B2E_MAX_LIMROWS) { sm.lim_d[slot] = lane; sm.lim_dist[slot] = dlo; }
    }
    if (up_hit) {
      const int slot = __popc(blo & lt) + __popc(bup & lt) + (lo_hit ? 1 : 0);
      if (slot < B2E_MAX_LIMROWS) { sm.lim_d[slot] = lane | (1 << 8); sm.lim_dist[slot] = dup; }
    }
    gsync(g);

    PROF_T(3);   // collision done
    PHASE_BARRIER();
    // ---- rows + PGS ----
    R = nd + nlim + 3 * nc;
    // generic rows (limits + contacts) come in sets of 16; both envs of the warp take the same
    // instantiation (the larger need) so the warp does not execute two instantiations back to back
    const int RG = nlim + 3 * nc;
    const int RGw = (int)__reduce_max_sync(FULL, (unsigned)RG);
    // storage of a big system (> 16 generic rows): an overflow slot of the block if one is free, else global scratch
    float* gscratch = st.scratch + (size_t)env * SCRATCH_PER_ENV;
    // (the claim loop is warp-uniform: a lane looping on its own over the slots left the warp diverged for the whole build --
    // every shuffle of it through the divergent path -- until the first ballot of the sweeps; active-lane probes, DESIGN.md 4c')
    int got = -1;
#ifdef PROFILE_WARM
    const int amB = __popc(__activemask());
#endif
    for (int k = 0; k < nslot; k++) {
      const bool want = lane == 0 && RG > GL && got < 0;
      if (!__any_sync(FULL, want)) break;
      if (want && atomicCAS(&slot_owner[k], -1, warp * 2 + half) == -1) got = k;
    }
#ifdef PROFILE_WARM
    const int amC = __popc(__activemask());
#endif
    got = SHF(got, 0);
    const bool need_global = __any_sync(FULL, RG > GL && got < 0);   // rare: more big systems in the block than slots
    const float* big = (RG > GL && got < 0) ? gscratch : reinterpret_cast<const float*>(&slots[got < 0 ? 0 : got]);
    PROF_T(4);   // past the pre-solve barrier
#ifdef PROFILE_STAGES
    prof_slot = RG > GL ? got : -3;   // overflow slot of this environment's system (-1: global scratch, -3: no big system)
    prof_mate = __shfl_sync(FULL, env, 16 * (1 - half));   // the environment of the other half of the warp
    prof_glob = need_global ? 1 : 0;
    prof_rgw = RGw;
#endif
#ifdef PROFILE_CYCLES
    const long long t_solve0 = clock64();
#endif
    const int big_off = got < 0 ? -1 : slots_off + (int)sizeof(BigSlot) * got;
#define B2E_SOLVE(N, G) build_and_solve<N, G>(sm, M, U, P, g.hm, g.sh, lane, nd, nlim, nc, my_q, my_target, my_kp, my_mv, cpos[0], cpos[1], cpos[2], big_off, gscratch)
    if (RGw <= GL) iters = B2E_SOLVE(1, false);
    else if (!need_global) iters = (RGw <= 2 * GL) ? B2E_SOLVE(2, false) : B2E_SOLVE(3, false);
    else iters = (RGw <= 2 * GL) ? B2E_SOLVE(2, true) : B2E_SOLVE(3, true);
#undef B2E_SOLVE
    PROF_T(5);   // solve done (this warp)
#ifdef PROFILE_WARM
    if (lane == 0) sm.cost[3] = (float)(amA + 100 * amB + 10000 * amC);   // active lanes after the collision stage / before / after the slot claim
#endif

#ifdef PROFILE_CYCLES
    R = (int)((clock64() - t_solve0) >> 6) * 64 + (RG > GL ? (got < 0 ? 2 : 1) : 0);   // cycles (multiple of 64) + storage code
    prof_pgs = (int)sm.lim_dist[3];   // cycles of the sweeps alone (replaces the contact count in B2E_F_STATUS)
#endif
    PHASE_BARRIER();
    PROF_T(6);   // past the post-solve barrier
    // ---- delta velocities dv = sum_r W_r * lambda_r (lane = velocity component) ----
    float dvk = 0.f;
    {
      const float* Wp = (RG <= GL) ? sm.W : big + BIGS * BIGS;
      for (int r = 0; r < RG; r++) dvk = fmaf(Wp[r * WSTRIDE + lane], sm.glam[r], dvk);
      // motor rows: W_d = column d of M^-1
      if (lane < NDMAX) {
#pragma unroll
        for (int d = 0; d < NDMAX; d++) dvk = fmaf(sm.Minv[lane][d], sm.mlam[d], dvk);
      }
    }
    // ---- integrate (semi-implicit Euler) ----
    if (is_dof && !ghost) {
      my_qd = vstar_d + dvk;
      my_q += dt * my_qd;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float dvl = SHF(dvk, NDMAX + k), dva = SHF(dvk, NDMAX + 3 + k);
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
      float dq[4] = {cw[0] * f, cw[1] * f, cw[2] * f, cs}, nq[4];
      quat_mul(dq, cquat, nq);
      const float nn = 1.0f / sqrtf(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
#pragma unroll
      for (int k = 0; k < 4; k++) cquat[k] = nq[k] * nn;
    }
    // ---- contact cache for the next step's warm start ----
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
      gsync(g);
      if (lane < B2E_CACHE_SLOTS && !ghost) {
        sm.ckey[lane] = key;
        sm.clam[lane][0] = l3[0]; sm.clam[lane][1] = l3[1]; sm.clam[lane][2] = l3[2];
      }
      gsync(g);
    }
    {
      bool bad = is_dof && !(isfinite(my_q) && isfinite(my_qd));
      bad = bad || !(isfinite(cpos[0]) && isfinite(cpos[1]) && isfinite(cpos[2]) && isfinite(cv[0]) && isfinite(cv[1]) &&
                     isfinite(cv[2]) && isfinite(cw[0]) && isfinite(cw[1]) && isfinite(cw[2]));
      if (gany(g, bad)) flags |= B2E_ST_NAN;
    }
    PROF_T(7);   // integrate + cache done

  }
  PROF_T(8);   // final FK + termination done

  if (role != ROLE_PLAIN) {   // file this environment under its cost class for the next full-batch step
    int cls = -1, rank = 0;
    if (lane == 0 && live_env) {
      const float c = sm.cost[0];
      cls = (R - nd > GL || c >= (float)tail_cost) ? NBK_MAIN : min(NBK_MAIN - 1, (int)(c * (1.0f / 8192.0f)));
      rank = atomicAdd(&cls_hist[cls], 1);
    }
    __syncthreads();
    if ((int)threadIdx.x < NBK) {
      const int h = cls_hist[threadIdx.x];
      cls_base[threadIdx.x] = h ? atomicAdd(&sched[SCHED_CNT(seq & 3, threadIdx.x)], h) : 0;
    }
    __syncthreads();
    if (cls >= 0) {
      const int B = st.B;
      int pos = cls_base[cls] + rank;
      if (cls == NBK_MAIN && pos >= tail_cap) {   // tail list full: the heaviest class of the main launch takes it
        cls = NBK_MAIN - 1;
        pos = atomicAdd(&sched[SCHED_CNT(seq & 3, cls)], 1);
      }
      sched[(cls == NBK_MAIN ? SCHED_TAIL(seq & 1, B) : SCHED_LIST(seq & 1, cls, B)) + pos] = env;
    }
  }
  PROF_T(9);   // filed under a cost class
  // ---- store state (padding groups shadow the last env: they keep pace but store nothing) ----
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

  PROF_T(10);  // state stored
  // ---- observation / termination / reward (panda_push_gym_env.py:249-253) ----
  if (mode == B2E_MODE_ACTION || obs_out) {
    const float* R2 = Rm;   // kinematics of the final q (computed at the top of the last loop pass)
    const float* p2 = pw;
    const int ee = U.ee_link;
    // EE COM pose
    float cm[3] = {U.ee_com[0], U.ee_com[1], U.ee_com[2]}, o[3];
    m3vec(R2, cm, o);
    float epos[3], Re[9];
#pragma unroll
    for (int k = 0; k < 3; k++) epos[k] = SHF(p2[k] + o[k], ee);
#pragma unroll
    for (int k = 0; k < 9; k++) Re[k] = SHF(R2[k], ee);
    // EE linear velocity: sum over the dofs on the path of (axis x (p_ee - o_j)) qd_j  /  axis qd_j
    float vl[3] = {0, 0, 0};
    {
      const unsigned eemask = U.ee_dofmask;
      const int jt = __ldg(&M->jtype[li]);
      float ax[3] = {__ldg(&M->axis[li][0]), __ldg(&M->axis[li][1]), __ldg(&M->axis[li][2])}, aw[3];
      m3vec(R2, ax, aw);
      const float qdl = SHF(my_qd, my_dof < 0 ? 0 : my_dof);
      if (lane < nl && my_dof >= 0 && ((eemask >> my_dof) & 1)) {
        if (jt == B2E_JOINT_REVOLUTE) {
          float rel[3] = {epos[0] - p2[0], epos[1] - p2[1], epos[2] - p2[2]}, t[3];
          cross3(aw, rel, t);
          vl[0] = t[0] * qdl; vl[1] = t[1] * qdl; vl[2] = t[2] * qdl;
        } else {
          vl[0] = aw[0] * qdl; vl[1] = aw[1] * qdl; vl[2] = aw[2] * qdl;
        }
      }
      // fixed-order sum over links 0..nl-1 so the result does not depend on lane scheduling
      float acc[3] = {0, 0, 0};
      for (int j = 0; j < nl; j++) {
        acc[0] += SHF(vl[0], j); acc[1] += SHF(vl[1], j); acc[2] += SHF(vl[2], j);
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
      float dd[3] = {cpos[0] - epos[0], cpos[1] - epos[1], cpos[2] - epos[2]}, Rh[9];
      quat_to_mat(hqi, Rh);
      m3vec(Rh, dd, relp);
    }
    quat_mul(hqi, oq, relq);
    quat_to_euler(relq, releu);
    float* obsb = sm.W;   // the W table is dead by now: staging for the observation vector
    gsync(g);
    if (lane == 0) {
      int n = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = epos[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = eeu[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = (vl[k] - P.vel_mean[k]) / P.vel_std[k];
      n += nd;
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = cpos[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = ceu[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = relp[k];
#pragma unroll
      for (int k = 0; k < 3; k++) obsb[n++] = releu[k];
      if (P.task == B2E_TASK_PUSH || P.task == B2E_TASK_GRASP) {
#pragma unroll
        for (int k = 0; k < 3; k++) obsb[n++] = target[k];
      }
    }
    if (is_dof) obsb[9 + lane] = my_q;
    gsync(g);
    for (int k = lane; k < P.n_obs; k += GL) {
      const float raw = obsb[k];
      if (!live_env) continue;
      st.raw_obs[(size_t)env * P.n_obs + k] = raw;
      if (obs_out) obs_out[(size_t)env * P.n_obs + k] = 2.0f * ((raw - P.obs_low[k]) / (P.obs_high[k] - P.obs_low[k])) - 1.0f;
    }
    float dd1[3] = {epos[0] - cpos[0], epos[1] - cpos[1], epos[2] - cpos[2]};
    const float d1 = sqrtf(dot3(dd1, dd1));
    float rew;
    int dn = 0;
    if (P.task == B2E_TASK_PUSH) {
      float dd2[3] = {cpos[0] - target[0], cpos[1] - target[1], cpos[2] - target[2]};
      const float d2 = sqrtf(dot3(dd2, dd2));
      if (P.goal_env) {
        dn = (counter > P.max_steps) || (d2 <= P.dist_min);
        rew = d2 > P.dist_min ? -1.0f : 0.0f;
      } else {
        if (d2 <= P.dist_min) { terminated = 1; dn = 1; }
        else if (terminated || counter > P.max_steps) dn = 1;
        rew = -d1 - d2;
        if (d2 <= P.dist_min) rew = 1000.0f + (100.0f - d2 * 80.0f);
      }
    } else if (P.task == B2E_TASK_GRASP) {
      const float lift = cpos[2] - P.grasp_rest_z;
      const bool success = lift >= P.grasp_lift;
      if (success) { terminated = 1; dn = 1; }
      else if (terminated || counter > P.max_steps) dn = 1;
      rew = -d1 + 50.0f * fminf(fmaxf(lift, 0.f), P.grasp_lift);
      if (success) rew = 1000.0f;
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
#ifdef PROFILE_STAGES
  PROF_T(11);   // observation / reward done
  if (lane == 0 && live_env && nsub > 0) {
    const long long t_end = clock64();
    float* o = st.contacts + (size_t)env * B2E_MAX_CONTACTS * 8;
    for (int k = 1; k < 7; k++) o[k - 1] = (float)(prof_t[k] - prof_t[0]);
    o[6] = (float)(t_end - prof_t[0]);
    o[7] = (float)blockIdx.x;
    for (int k = 7; k < 11; k++) o[13 + k - 7] = (float)(prof_t[k] - prof_t[0]);   // [13..16]: finer stamps of the `rest`
    o[17] = sm.cost[1]; o[18] = sm.cost[2]; o[19] = sm.cost[3];                    // [17..19]: parts of the solve (cycles)
    o[21] = (float)(role * 100 + prof_slot);                                        // [21]: launch role x 100 + overflow slot
    o[22] = (float)prof_mate; o[23] = (float)(prof_glob * 1000 + prof_rgw); o[24] = (float)n_slots;
#ifdef PROFILE_WARM
    for (int k = 0; k < 14; k++) o[32 + k] = (float)prof_am[k];
#endif
#ifdef PROFILE_SWEEP
    for (int k = 0; k < 5; k++) o[8 + k] = st.scratch[(size_t)env * SCRATCH_PER_ENV + SCRATCH_PER_ENV - 8 + k];   // [8..12]: parts of the sweeps
    o[20] = st.scratch[(size_t)env * SCRATCH_PER_ENV + SCRATCH_PER_ENV - 8 + 5];                                   // [20]: before the serial motor rows
#endif
  }
#endif
  if (lane == 0 && mode != B2E_MODE_OBSERVE && live_env) {
    st.counters[env * 2] = counter;
    st.counters[env * 2 + 1] = terminated;
    st.status[env * 4 + 0] = flags;
    if (nsub > 0) {   // solver diagnostics of the last physics step survive pure observation calls
      st.status[env * 4 + 1] = iters;
#ifdef PROFILE_CYCLES
      st.status[env * 4 + 2] = prof_pgs;
#else
      st.status[env * 4 + 2] = nc;
#endif
      st.status[env * 4 + 3] = R;
    }
  }
}

// The tree kernel (iCub) keeps the closed-form contact set (cube vertices vs table top / ground plane, spheres vs cube /
// table top): its arms cannot reach the table rim, the legs or the floor in the pinned pose; box-box families, static
// boxes and self pairs are group-kernel (Panda) features.
#define KEY_CUBE_TABLE B2E_KEY_CUBE_TABLE
#define KEY_CUBE_PLANE B2E_KEY_CUBE_PLANE
#define KEY_SPHERE_CUBE B2E_KEY_SPHERE_CUBE
#define KEY_SPHERE_TABLE B2E_KEY_SPHERE_TABLE
#define CT_SPHERE_CUBE CT_ARM_CUBE
#define CT_SPHERE_STATIC CT_ARM_STATIC
#include "b2env_tree.cuh"

// masked reset: home joint state, object pose, target, counters, empty cache
__global__ void reset_kernel(const DevModel* __restrict__ M, const b2e_params P, DevState st,
                             const uint8_t* __restrict__ mask, const float* __restrict__ obj_init_pose,
                             const float* __restrict__ target) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= st.B) return;
  if (mask && !mask[env]) return;
  const int nd = M->n_dof;
  for (int d = 0; d < nd; d++) {
    st.q[env * nd + d] = M->home[d];
    st.qd[env * nd + d] = 0.f;
    st.mtarget[env * nd + d] = M->home[d];
  }
  for (int k = 0; k < 7; k++) st.obj_pose[env * 7 + k] = obj_init_pose[env * 7 + k];
  for (int k = 0; k < 6; k++) st.obj_vel[env * 6 + k] = 0.f;
  for (int k = 0; k < 3; k++) st.target[env * 3 + k] = target[env * 3 + k];
  st.counters[env * 2] = 0; st.counters[env * 2 + 1] = 0;
  for (int s = 0; s < B2E_CACHE_SLOTS; s++) {
    st.cache_key[env * B2E_CACHE_SLOTS + s] = -1;
    for (int j = 0; j < 3; j++) st.cache_lam[(env * B2E_CACHE_SLOTS + s) * 3 + j] = 0.f;
  }
  for (int k = 0; k < 6; k++) st.hand_pose[env * 6 + k] = P.home_hand_pose[k];
  for (int k = 0; k < 4; k++) st.status[env * 4 + k] = 0;
  st.shaping[env * 2] = 1.f; st.shaping[env * 2 + 1] = 1.f;
}

// ------------------------------------------------------------------------------------------
// host side: C-ABI
static int field_width(const b2e_sim* s, int f) {
  const int nd = s->model.n_dof;
  switch (f) {
    case B2E_F_Q: case B2E_F_QD: case B2E_F_MTARGET: return nd;
    case B2E_F_OBJ_POSE: return 7;
    case B2E_F_OBJ_VEL: return 6;
    case B2E_F_TARGET: return 3;
    case B2E_F_COUNTERS: return 2;
    case B2E_F_CACHE_KEY: return B2E_CACHE_SLOTS;
    case B2E_F_CACHE_LAM: return B2E_CACHE_SLOTS * 3;
    case B2E_F_HAND_POSE: return 6;
    case B2E_F_STATUS: return 4;
    case B2E_F_RAW_OBS: return s->params.n_obs;
    case B2E_F_CONTACTS: return B2E_MAX_CONTACTS * 8;
    case B2E_F_SHAPING: return 2;
    default: return -1;
  }
}

static int build_dev_model(const b2e_model* m, DevModel* d, DevModelU* u, int* tree_out) {
  memset(d, 0, sizeof(*d));
  memset(u, 0, sizeof(*u));
  if (m->n_links < 1 || m->n_dof < 1) return fail(B2E_EINVAL, "empty model%s", "");
  // kernel choice: the 16-lane group kernel for small chains (the Panda), else the warp-per-env tree kernel,
  // which wants one body per movable joint (model.merge_fixed_links) so that lane = body = dof
  const bool tree = m->n_links > TLMAX || m->n_dof > NDMAX;
  *tree_out = tree ? 1 : 0;
  if (tree) {
    if (m->n_links > NLMAX || m->n_dof != m->n_links)
      return fail(B2E_EUNSUPPORTED, "tree kernel: <= 32 bodies, one per movable joint (merge fixed links on the host)%s", "");
    for (int i = 0; i < m->n_links; i++)
      if (m->dof[i] != i || m->jtype[i] == B2E_JOINT_FIXED)
        return fail(B2E_EUNSUPPORTED, "tree kernel: body i must carry dof i%s", "");
  }
  if (m->n_spheres > B2E_MAX_SPHERES || m->n_spheres < 0) return fail(B2E_EINVAL, "bad n_spheres%s", "");
  d->n_links = m->n_links; d->n_dof = m->n_dof; d->ee_link = m->ee_link; d->n_spheres = m->n_spheres;
  int maxdepth = 1;
  for (int i = 0; i < m->n_links; i++) {
    if (m->parent[i] >= i) return fail(B2E_EINVAL, "links must be ordered parent-first%s", "");
    d->parent[i] = m->parent[i]; d->jtype[i] = m->jtype[i]; d->dof[i] = m->dof[i];
    if (m->dof[i] >= 0) d->dof_link[m->dof[i]] = i;
    for (int k = 0; k < 3; k++) { d->jpos[i][k] = m->jpos[i][k]; d->axis[i][k] = m->axis[i][k]; d->com[i][k] = m->com[i][k]; }
    for (int k = 0; k < 9; k++) { d->jrot[i][k] = m->jrot[i][k]; d->inertia[i][k] = m->inertia[i][k]; }
    d->mass[i] = m->mass[i];
    unsigned mask = 0;
    int depth = 0;
    for (int j = i; j >= 0; j = m->parent[j]) {
      if (m->dof[j] >= 0) mask |= 1u << m->dof[j];
      depth++;
    }
    d->link_dofmask[i] = mask;
    if (depth > maxdepth) maxdepth = depth;
  }
  int rounds = 0;
  while ((1 << rounds) < maxdepth) rounds++;
  d->fk_rounds = rounds;
  if (tree) {
    // child -> parent accumulation schedule: every round, each body pulls at most one child whose own subtree is
    // complete (greedy, deepest bodies first)
    const int n = m->n_links;
    int pending[NLMAX], folded[NLMAX];
    for (int i = 0; i < n; i++) { pending[i] = 0; folded[i] = 0; }
    for (int i = 0; i < n; i++) if (m->parent[i] >= 0) pending[m->parent[i]]++;
    int left = 0, rounds = 0;
    for (int i = 0; i < n; i++) if (m->parent[i] >= 0) left++;
    while (left > 0) {
      if (rounds >= 24) return fail(B2E_EUNSUPPORTED, "kinematic tree needs more than 24 accumulation rounds%s", "");
      int busy[NLMAX], pick[NLMAX], np = 0;
      for (int l = 0; l < NLMAX; l++) { d->tsched[rounds][l] = -1; busy[l] = 0; }
      for (int i = n - 1; i >= 0; i--) {
        const int p = m->parent[i];
        if (p < 0 || folded[i] || pending[i] > 0 || busy[p]) continue;
        d->tsched[rounds][p] = (signed char)i;
        busy[p] = 1;
        pick[np++] = i;
      }
      for (int k = 0; k < np; k++) { folded[pick[k]] = 1; pending[m->parent[pick[k]]]--; left--; }
      rounds++;
    }
    u->acc_rounds = rounds;
  } else {
    // child->parent accumulation schedule (heavy-path decomposition)
    const int n = m->n_links;
    int size[TLMAX], heavy[TLMAX], depth[TLMAX], head[TLMAX];
    for (int i = 0; i < n; i++) { size[i] = 1; heavy[i] = -1; }
    for (int i = n - 1; i >= 0; i--) if (m->parent[i] >= 0) size[m->parent[i]] += size[i];
    for (int i = 0; i < n; i++) {
      depth[i] = m->parent[i] < 0 ? 0 : depth[m->parent[i]] + 1;
      int p = m->parent[i];
      if (p >= 0 && (heavy[p] < 0 || size[i] > size[heavy[p]])) heavy[p] = i;
    }
    for (int i = 0; i < n; i++) head[i] = (m->parent[i] >= 0 && heavy[m->parent[i]] == i) ? head[m->parent[i]] : i;
    int src[16][NLMAX];
    for (int r = 0; r < 16; r++) for (int l = 0; l < NLMAX; l++) src[r][l] = 15;
    int rounds = 0;
    // path heads, deepest first
    int heads[TLMAX], nh = 0;
    for (int i = 0; i < n; i++) if (head[i] == i) heads[nh++] = i;
    for (int a = 0; a < nh; a++) for (int b = a + 1; b < nh; b++) if (depth[heads[b]] > depth[heads[a]]) { int t = heads[a]; heads[a] = heads[b]; heads[b] = t; }
    for (int a = 0; a < nh; a++) {
      int path[TLMAX], len = 0;
      for (int v = heads[a]; v >= 0; v = heavy[v]) path[len++] = v;
      for (int step = 1; step < len; step <<= 1) {
        if (rounds >= 16) return fail(B2E_EUNSUPPORTED, "kinematic tree needs more than 16 accumulation rounds%s", "");
        for (int i = 0; i + step < len; i++) src[rounds][path[i]] = path[i + step];
        rounds++;
      }
      if (m->parent[heads[a]] >= 0) {
        if (rounds >= 16) return fail(B2E_EUNSUPPORTED, "kinematic tree needs more than 16 accumulation rounds%s", "");
        src[rounds][m->parent[heads[a]]] = heads[a];
        rounds++;
      }
    }
    u->acc_rounds = rounds;
    for (int l = 0; l < NLMAX; l++) {
      unsigned w0 = 0, w1 = 0;
      for (int r = 0; r < 8; r++) { w0 |= (unsigned)src[r][l] << (4 * r); w1 |= (unsigned)src[8 + r][l] << (4 * r); }
      d->acc_sched[0][l] = w0; d->acc_sched[1][l] = w1;
    }
  }
  u->n_links = m->n_links; u->n_dof = m->n_dof; u->ee_link = m->ee_link; u->n_spheres = m->n_spheres;
  u->fk_rounds = d->fk_rounds; u->ee_dofmask = d->link_dofmask[m->ee_link];
  for (int k = 0; k < 3; k++) { u->base_pos[k] = m->base_pos[k]; u->ee_com[k] = m->com[m->ee_link][k]; }
  for (int k = 0; k < 9; k++) u->base_rot[k] = m->base_rot[k];
  for (int i = 0; i < TLMAX; i++) u->parent[i] = i < m->n_links ? m->parent[i] : -1;

  for (int k = 0; k < m->n_dof; k++) {
    d->lower[k] = m->lower[k]; d->upper[k] = m->upper[k]; d->limit_margin[k] = m->limit_margin[k];
    d->max_force[k] = m->max_force[k]; d->max_vel[k] = m->max_vel[k]; d->joint_damping[k] = m->joint_damping[k];
    d->home[k] = m->home[k];
  }
  for (int k = 0; k < 3; k++) d->base_pos[k] = m->base_pos[k];
  for (int k = 0; k < 9; k++) d->base_rot[k] = m->base_rot[k];
  for (int s = 0; s < m->n_spheres; s++) {
    d->sph_link[s] = m->sph_link[s];
    for (int k = 0; k < 3; k++) d->sph_c[s][k] = m->sph_c[s][k];
    d->sph_r[s] = m->sph_r[s]; d->sph_mu[s] = m->sph_mu[s]; d->sph_erp[s] = m->sph_erp[s]; d->sph_cfm[s] = m->sph_cfm[s];
  }
  if (m->n_boxes < 0 || m->n_boxes > B2E_MAX_BOXES || m->n_self_pairs < 0 || m->n_self_pairs > B2E_MAX_SELF_PAIRS)
    return fail(B2E_EINVAL, "bad n_boxes / n_self_pairs%s", "");
  if (!tree && m->n_boxes > 2) return fail(B2E_EUNSUPPORTED, "group kernel: at most two box proxies%s", "");
  if (tree) {   // the tree kernel collides sphere proxies only: other proxies are refused, not dropped (the oracle would use them)
    bool flagged = false;
    for (int s = 0; s < m->n_spheres; s++) flagged = flagged || m->sph_flags[s] != 0;
    if (m->n_boxes > 0 || m->n_self_pairs > 0 || m->n_caps > 0 || flagged)
      return fail(B2E_EUNSUPPORTED, "tree kernel: box / capsule proxies, self-collision pairs and sphere flags need the group kernel%s", "");
  }
  d->n_boxes = u->n_boxes = tree ? 0 : m->n_boxes;
  d->n_self_pairs = u->n_self_pairs = tree ? 0 : m->n_self_pairs;
  for (int b = 0; b < m->n_boxes; b++) {
    d->box_link[b] = m->box_link[b];
    for (int k = 0; k < 3; k++) { d->box_c[b][k] = m->box_c[b][k]; d->box_h[b][k] = m->box_h[b][k]; }
    d->box_mu[b] = m->box_mu[b]; d->box_erp[b] = m->box_erp[b]; d->box_cfm[b] = m->box_cfm[b];
  }
  for (int k = 0; k < m->n_self_pairs; k++) {
    if (m->self_a[k] < 0 || m->self_a[k] >= m->n_spheres || m->self_b[k] < 0 || m->self_b[k] >= m->n_spheres)
      return fail(B2E_EINVAL, "self-collision pair names a sphere that does not exist%s", "");
    d->self_a[k] = m->self_a[k]; d->self_b[k] = m->self_b[k];
  }
  for (int s = 0; s < m->n_spheres; s++) d->sph_flags[s] = tree ? 0 : m->sph_flags[s];
  if (m->n_caps < 0 || m->n_caps > B2E_MAX_CAPS) return fail(B2E_EINVAL, "bad n_caps%s", "");
  d->n_caps = u->n_caps = tree ? 0 : m->n_caps;
  for (int c = 0; c < m->n_caps; c++) {
    if (m->cap_link[c] < 0 || m->cap_link[c] >= m->n_links) return fail(B2E_EINVAL, "capsule on a link that does not exist%s", "");
    d->cap_link[c] = m->cap_link[c];
    for (int k = 0; k < 3; k++) { d->cap_p0[c][k] = m->cap_p0[c][k]; d->cap_p1[c][k] = m->cap_p1[c][k]; }
    d->cap_r[c] = m->cap_r[c]; d->cap_mu[c] = m->cap_mu[c];
  }
  return 0;
}

#define SMEM_BYTES_FOR(wpb, nslot) (sizeof(EnvSmem) * 2 * (wpb) + sizeof(BigSlot) * (nslot) + (16 + 2 * NBK + 2) * sizeof(int))
#define SMEM_BYTES SMEM_BYTES_FOR(WPB, NSLOT)
// 228 KB per SM, 1 KB reserved per resident block: a tail block of this size leaves < SMEM_BYTES + 1 KB free
#define TAIL_EXCL_SMEM (228 * 1024 - 2 * 1024 - SMEM_BYTES + 512)
#define TREE_SMEM_BYTES (sizeof(TreeSmem) * TREE_WPB)
template <bool IK>
static void launch_group(b2e_sim* s, int blocks, int threads, size_t smem, void* stream, const float* action, float* obs, float* reward,
                         float* done, int n_substeps, int mode, const int* env_ids, int n, int env_offset, int* sched, int seq, int role,
                         int nslot) {
  B2E_LAUNCH(step_kernel<IK>, blocks, threads, smem, stream, s->d_model, s->umodel, s->params, s->st, action, obs, reward, done,
             n_substeps, mode, s->record_contacts, env_ids, n, env_offset, sched, seq, role, nslot,
             s->tail_blocks * s->tail_wpb * 2, s->tail_cost);
}
static int launch_step(b2e_sim* s, const float* action, float* obs, float* reward, float* done, int n_substeps, int mode,
                       const int* env_ids, int n_ids, void* stream, int env_offset = 0, bool ordered = true) {
  const int n = (env_ids || n_ids > 0) ? n_ids : s->B;
  if (n <= 0) return 0;
  if (ordered) ORDER_BEGIN(s, stream);
  if (s->tree) {
    const int tb = (n + TREE_WPB - 1) / TREE_WPB;
    // cost-ordered blocks only for physics launches over the whole batch
    int* tsched = (!env_ids && env_offset == 0 && n == s->B && n_substeps > 0 && s->d_sched) ? s->d_sched : nullptr;
    const int tseq = tsched ? s->sched_seq++ : 0;
    if (s->params.use_ik)
      B2E_LAUNCH(tree_step_kernel<true>, tb, 32 * TREE_WPB, TREE_SMEM_BYTES, stream,
          s->d_model, s->umodel, s->params, s->st, action, obs, reward, done, n_substeps, mode, s->record_contacts, env_ids, n,
          env_offset, tsched, tseq);
    else
      B2E_LAUNCH(tree_step_kernel<false>, tb, 32 * TREE_WPB, TREE_SMEM_BYTES, stream,
          s->d_model, s->umodel, s->params, s->st, action, obs, reward, done, n_substeps, mode, s->record_contacts, env_ids, n,
          env_offset, tsched, tseq);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    if (ordered) ORDER_END(s, stream);
    return 0;
  }
  // cost-ordered scheduling only for physics launches over the whole batch
  int* sched = (!env_ids && env_offset == 0 && n == s->B && n_substeps > 0 && s->d_sched) ? s->d_sched : nullptr;
  const bool ik = s->params.use_ik != 0;
  if (!sched) {
    const int blocks = (n + 2 * WPB - 1) / (2 * WPB);
    if (ik) launch_group<true>(s, blocks, 32 * WPB, SMEM_BYTES, stream, action, obs, reward, done, n_substeps, mode, env_ids, n, env_offset, nullptr, 0, ROLE_PLAIN, NSLOT);
    else launch_group<false>(s, blocks, 32 * WPB, SMEM_BYTES, stream, action, obs, reward, done, n_substeps, mode, env_ids, n, env_offset, nullptr, 0, ROLE_PLAIN, NSLOT);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    if (ordered) ORDER_END(s, stream);
    return 0;
  }
  const int seq = s->sched_seq++;
  // fork: the tail launch (heavy class, lean blocks, own high-priority stream) is enqueued FIRST so that its blocks are
  // resident when the main launch starts filling the machine; join: the caller's stream waits for it
  const int twpb = s->tail_wpb, tslots = 2 * twpb;
  const int tblocks = s->tail_blocks;
  // shared memory of a tail block: what it uses, or (tail_excl) enough that no block of the main launch fits next to it — the
  // few environments that decide how long the step lasts then run without other warps competing for issue slots and for
  // the shared-memory / shuffle pipe (tools/micro/row_chain3.cu: 56 cycles per row update alone, 125 next to 16 busy warps)
  size_t tsmem = SMEM_BYTES_FOR(twpb, tslots);
  if (s->tail_excl && tsmem < (size_t)TAIL_EXCL_SMEM) tsmem = TAIL_EXCL_SMEM;
#ifndef B2E_EMU
  CUDA_TRY(cudaEventRecord(s->ev_fork, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamWaitEvent(s->tstream, s->ev_fork, 0));
#endif
  void* ts = (void*)s->tstream;
  if (ik) launch_group<true>(s, tblocks, 32 * twpb, tsmem, ts, action, obs, reward, done, n_substeps, mode, nullptr, n, 0, sched, seq, ROLE_TAIL, tslots);
  else launch_group<false>(s, tblocks, 32 * twpb, tsmem, ts, action, obs, reward, done, n_substeps, mode, nullptr, n, 0, sched, seq, ROLE_TAIL, tslots);
  CUDA_TRY(cudaGetLastError());
#ifndef B2E_EMU
  CUDA_TRY(cudaEventRecord(s->ev_join, s->tstream));
#endif
  const int blocks = (n + 2 * WPB - 1) / (2 * WPB);
  if (ik) launch_group<true>(s, blocks, 32 * WPB, SMEM_BYTES, stream, action, obs, reward, done, n_substeps, mode, nullptr, n, 0, sched, seq, ROLE_MAIN, NSLOT);
  else launch_group<false>(s, blocks, 32 * WPB, SMEM_BYTES, stream, action, obs, reward, done, n_substeps, mode, nullptr, n, 0, sched, seq, ROLE_MAIN, NSLOT);
  CUDA_TRY(cudaGetLastError());
#ifndef B2E_EMU
  CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s->ev_join, 0));
#endif
  s->launches += 2;
  if (ordered) ORDER_END(s, stream);
  return 0;
}

// scatter / gather rows of a state field
__global__ void rows_kernel(float* field, const int* __restrict__ ids, int n, int width, float* rows, int gather) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int r = i / width, c = i - r * width;
  if (gather) rows[i] = field[(size_t)ids[r] * width + c];
  else field[(size_t)ids[r] * width + c] = rows[i];
}

extern "C" {

const char* b2e_last_error(void) { return g_err; }
#ifdef B2E_EMU
const char* b2e_version(void) { return "b2env 0.4 EMULATION (host build of the CUDA source: test infrastructure, not the product)"; }
#else
const char* b2e_version(void) { return "b2env 0.4 (sm_100a; Panda: two environments per warp, slow-first block order; iCub: one warp per environment, phase-locked blocks)"; }
#endif

int b2e_field_elem_size(int field) {
  (void)field;
  return 4;
}
int b2e_field_width(const b2e_sim* sim, int field) { return sim ? field_width(sim, field) : -1; }
int b2e_num_envs(const b2e_sim* sim) { return sim ? sim->B : -1; }
int64_t b2e_launch_count(const b2e_sim* sim) { return sim ? sim->launches : -1; }

static int create_impl(b2e_sim* s, const DevModel& hm, const b2e_params* params, int num_envs, int tree);

// Parameters one of the two kernels does not implement are refused, never ignored (the oracle is generic; a silent difference
// would be a parity hole): joint lists / iCub rewards / the hand-frame IK offset need the tree kernel, the grasp task needs the
// Panda kernel.
static int check_params(const char* who, const b2e_params* p, int tree) {
  if (!tree) {
    if (p->n_obs_joints > 0) return fail(B2E_EUNSUPPORTED, "%s: joint lists (n_obs_joints > 0) need the tree kernel", who);
    if (p->reward_kind != B2E_REWARD_PANDA) return fail(B2E_EUNSUPPORTED, "%s: reward_kind != B2E_REWARD_PANDA needs the tree kernel", who);
    if (p->ik_link_offset[0] != 0.f || p->ik_link_offset[1] != 0.f || p->ik_link_offset[2] != 0.f)
      return fail(B2E_EUNSUPPORTED, "%s: ik_link_offset needs the tree kernel", who);
  } else {
    if (p->task == B2E_TASK_GRASP) return fail(B2E_EUNSUPPORTED, "%s: the grasp task needs the Panda model", who);
  }
  if (p->n_sboxes < 0 || p->n_sboxes > B2E_MAX_SBOXES) return fail(B2E_EINVAL, "%s: bad n_sboxes", who);
  return 0;
}

int b2e_create(const b2e_model* model, const b2e_params* params, int num_envs, int device, b2e_sim** out) {
  if (!model || !params || !out || num_envs < 1) return fail(B2E_EINVAL, "b2e_create: bad argument%s", "");
  if (params->n_obs > B2E_MAX_OBS || params->n_obs < 1) return fail(B2E_EINVAL, "b2e_create: bad n_obs%s", "");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(B2E_ECUDA, "b2e_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(B2E_EINVAL, "b2e_create: bad device index%s", "");
  CUDA_TRY(cudaSetDevice(device));
  DevModel hm;
  DevModelU hu;
  int tree = 0;
  int rc = build_dev_model(model, &hm, &hu, &tree);
  if (rc) return rc;
  rc = check_params("b2e_create", params, tree);
  if (rc) return rc;
  b2e_sim* s = new b2e_sim();
  memset(s, 0, sizeof(*s));
  s->B = num_envs; s->device = device; s->model = *model; s->params = *params; s->umodel = hu; s->tree = tree;
  rc = create_impl(s, hm, params, num_envs, tree);
  if (rc) {   // nothing allocated so far survives a failed create (g_err keeps the first error)
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    b2e_destroy(s);
    memcpy(g_err, keep, sizeof(keep));
    return rc;
  }
  *out = s;
  return 0;
}

static int create_impl(b2e_sim* s, const DevModel& hm, const b2e_params* params, int num_envs, int tree) {
  CUDA_TRY(cudaMalloc(&s->d_model, sizeof(DevModel)));
  CUDA_TRY(cudaMemcpy(s->d_model, &hm, sizeof(DevModel), cudaMemcpyHostToDevice));
  for (int f = 0; f < B2E_F_COUNT; f++) {
    size_t bytes = (size_t)num_envs * field_width(s, f) * 4;
    CUDA_TRY(cudaMalloc(&s->fields[f], bytes));
    CUDA_TRY(cudaMemset(s->fields[f], 0, bytes));
  }
  CUDA_TRY(cudaMemset(s->fields[B2E_F_CACHE_KEY], 0xff, (size_t)num_envs * B2E_CACHE_SLOTS * 4));
  s->st.B = num_envs;
  s->st.q = (float*)s->fields[B2E_F_Q]; s->st.qd = (float*)s->fields[B2E_F_QD];
  s->st.obj_pose = (float*)s->fields[B2E_F_OBJ_POSE]; s->st.obj_vel = (float*)s->fields[B2E_F_OBJ_VEL];
  s->st.target = (float*)s->fields[B2E_F_TARGET]; s->st.mtarget = (float*)s->fields[B2E_F_MTARGET];
  s->st.counters = (int*)s->fields[B2E_F_COUNTERS]; s->st.cache_key = (int*)s->fields[B2E_F_CACHE_KEY];
  s->st.cache_lam = (float*)s->fields[B2E_F_CACHE_LAM]; s->st.hand_pose = (float*)s->fields[B2E_F_HAND_POSE];
  s->st.status = (int*)s->fields[B2E_F_STATUS]; s->st.raw_obs = (float*)s->fields[B2E_F_RAW_OBS];
  s->st.contacts = (float*)s->fields[B2E_F_CONTACTS];
  s->st.shaping = (float*)s->fields[B2E_F_SHAPING];
  // + SCRATCH_PAD floats: the lanes of the third row set read W^T up to 7 entries past a system's end (BigSlot carries the same
  // pad); for every environment but the last that is the neighbour's scratch, for the last one it must still be allocated
  // (found by compute-sanitizer on a 70-env batch whose last block had more big systems than overflow slots)
  CUDA_TRY(cudaMalloc(&s->st.scratch, ((size_t)num_envs * SCRATCH_PER_ENV + SCRATCH_PAD) * 4));
  CUDA_TRY(cudaMemset(s->st.scratch, 0, ((size_t)num_envs * SCRATCH_PER_ENV + SCRATCH_PAD) * 4));   // tables are finite from the start
  {
    const char* e = getenv("B2ENV_SCHED");   // B2ENV_SCHED=0 switches cost-ordered scheduling off (A/B measurements)
    const char* mn = getenv("B2ENV_SCHED_MIN");   // smallest batch that is scheduled (tests lower it)
    if (num_envs >= (mn ? atoi(mn) : 2048) && !(e && e[0] == '0')) {   // group kernel: class lists + tail launch; tree kernel: class lists
      // every environment starts in class 0, in index order (as the lists of "step 1", read by the first step, seq = 2)
      const size_t ni = (size_t)SCHED_INTS(num_envs);
      int* h = (int*)calloc(ni, sizeof(int));
      if (!h) return fail(B2E_EINVAL, "b2e_create: out of host memory%s", "");
      h[SCHED_CNT(1, 0)] = num_envs;
      for (int i = 0; i < num_envs; i++) h[SCHED_LIST(1, 0, num_envs) + i] = i;
      cudaError_t e1 = cudaMalloc(&s->d_sched, ni * sizeof(int));
      if (e1 == cudaSuccess) e1 = cudaMemcpy(s->d_sched, h, ni * sizeof(int), cudaMemcpyHostToDevice);
      free(h);
      CUDA_TRY(e1);
      s->sched_seq = 2;
      const char* w = getenv("B2ENV_TAIL_WPB");   // warps per block of the tail launch (1, 2 or 4)
      s->tail_wpb = (w && (w[0] == '1' || w[0] == '2' || w[0] == '4')) ? (w[0] - '0') : 4;
      const char* tb = getenv("B2ENV_TAIL_BLOCKS");
      s->tail_blocks = tb ? atoi(tb) : 64;
      if (s->tail_blocks < 1) s->tail_blocks = 1;
      if (s->tail_blocks * s->tail_wpb * 2 > TAIL_CAP) s->tail_blocks = TAIL_CAP / (s->tail_wpb * 2);
      const char* te = getenv("B2ENV_TAIL_EXCL");   // 0: tail blocks share their SM with the main launch (A/B runs)
      s->tail_excl = !(te && te[0] == '0');
      const char* tc = getenv("B2ENV_TAIL_COST");
      s->tail_cost = tc ? atoi(tc) : SCHED_TAIL_COST;
#ifndef B2E_EMU
      int lo_p = 0, hi_p = 0;
      CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
      CUDA_TRY(cudaStreamCreateWithPriority(&s->tstream, cudaStreamNonBlocking, hi_p));
      CUDA_TRY(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
#endif
    }
  }
  const size_t na = (size_t)num_envs * (params->n_act > 0 ? params->n_act : 1) * 4, no = (size_t)num_envs * params->n_obs * 4;
  CUDA_TRY(cudaMalloc(&s->d_action, na)); CUDA_TRY(cudaMalloc(&s->d_obs, no));
  CUDA_TRY(cudaMalloc(&s->d_reward, (size_t)num_envs * 4)); CUDA_TRY(cudaMalloc(&s->d_done, (size_t)num_envs * 4));
  CUDA_TRY(cudaMallocHost(&s->h_action, na)); CUDA_TRY(cudaMallocHost(&s->h_obs, no));
  CUDA_TRY(cudaMallocHost(&s->h_reward, (size_t)num_envs * 4)); CUDA_TRY(cudaMallocHost(&s->h_done, (size_t)num_envs * 4));
  CUDA_TRY(cudaEventCreate(&s->ev0)); CUDA_TRY(cudaEventCreate(&s->ev1));
  CUDA_TRY(cudaEventCreateWithFlags(&s->ev_order, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(s->ev_order, 0));
  CUDA_TRY(cudaStreamCreate(&s->pstream[0])); CUDA_TRY(cudaStreamCreate(&s->pstream[1]));
  int smem_max = (int)(SMEM_BYTES > SMEM_BYTES_FOR(4, 8) ? SMEM_BYTES : SMEM_BYTES_FOR(4, 8));   // main / largest tail block
  if (smem_max < (int)TAIL_EXCL_SMEM) smem_max = (int)TAIL_EXCL_SMEM;                             // SM-exclusive tail block
  static_assert(TAIL_EXCL_SMEM <= 227 * 1024 && SMEM_BYTES_FOR(4, 8) <= 227 * 1024, "tail block within the 227 KB a block can have");
  CUDA_TRY(cudaFuncSetAttribute(step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  CUDA_TRY(cudaFuncSetAttribute(step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  CUDA_TRY(cudaFuncSetAttribute(tree_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TREE_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(tree_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TREE_SMEM_BYTES));
  return 0;
}

void b2e_destroy(b2e_sim* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (int f = 0; f < B2E_F_COUNT; f++) cudaFree(s->fields[f]);
  cudaFree(s->st.scratch);
  cudaFree(s->d_sched);
  cudaFree(s->d_model); cudaFree(s->d_action); cudaFree(s->d_obs); cudaFree(s->d_reward); cudaFree(s->d_done);
  cudaFreeHost(s->h_action); cudaFreeHost(s->h_obs); cudaFreeHost(s->h_reward); cudaFreeHost(s->h_done);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->ev_order) cudaEventDestroy(s->ev_order);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  if (s->tstream) cudaStreamDestroy(s->tstream);
  if (s->pstream[0]) cudaStreamDestroy(s->pstream[0]);
  if (s->pstream[1]) cudaStreamDestroy(s->pstream[1]);
  cudaGetLastError();   // a partially created sim frees null handles: leave no sticky error behind
  delete s;
}

int b2e_set_params(b2e_sim* s, const b2e_params* p) {
  if (!s || !p) return fail(B2E_EINVAL, "b2e_set_params: null%s", "");
  if (p->n_obs != s->params.n_obs || p->n_act != s->params.n_act)
    return fail(B2E_EINVAL, "b2e_set_params: n_obs / n_act cannot change after create%s", "");
  const int rc = check_params("b2e_set_params", p, s->tree);
  if (rc) return rc;
  s->params = *p;
  return 0;
}

int b2e_set_option(b2e_sim* s, int option, int value) {
  if (!s) return fail(B2E_EINVAL, "b2e_set_option: null%s", "");
  if (option == 0) { s->record_contacts = value; return 0; }
  return fail(B2E_EINVAL, "b2e_set_option: unknown option%s", "");
}

int b2e_reset(b2e_sim* s, const uint8_t* env_mask, const float* obj_init_pose, const float* target, void* stream) {
  if (!s || !obj_init_pose || !target) return fail(B2E_EINVAL, "b2e_reset: null argument%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  const int tpb = 128;
  ORDER_BEGIN(s, stream);
  B2E_LAUNCH(reset_kernel, (s->B + tpb - 1) / tpb, tpb, 0, stream, s->d_model, s->params, s->st, env_mask, obj_init_pose,
             target);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  ORDER_END(s, stream);
  return 0;
}

int b2e_step(b2e_sim* s, const float* action, float* obs, float* reward, float* done, int n_substeps, int mode,
             void* stream) {
  if (!s) return fail(B2E_EINVAL, "b2e_step: null sim%s", "");
  if (mode == B2E_MODE_ACTION && !action) return fail(B2E_EINVAL, "b2e_step: action is required in ACTION mode%s", "");
  if (n_substeps < 0) return fail(B2E_EINVAL, "b2e_step: n_substeps < 0%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  return launch_step(s, action, obs, reward, done, n_substeps, mode, nullptr, s->B, stream, 0);
}

int b2e_step_subset(b2e_sim* s, const int32_t* env_ids, int n_ids, const float* action, float* obs, float* reward,
                    float* done, int n_substeps, int mode, void* stream) {
  if (!s || !env_ids || n_ids < 0 || n_ids > s->B) return fail(B2E_EINVAL, "b2e_step_subset: bad argument%s", "");
  if (mode == B2E_MODE_ACTION && !action) return fail(B2E_EINVAL, "b2e_step_subset: action is required in ACTION mode%s", "");
  if (n_substeps < 0) return fail(B2E_EINVAL, "b2e_step_subset: n_substeps < 0%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  return launch_step(s, action, obs, reward, done, n_substeps, mode, env_ids, n_ids, stream);
}

int b2e_set_rows(b2e_sim* s, int field, const int32_t* env_ids, int n, const void* rows, void* stream) {
  if (!s || !env_ids || !rows || field < 0 || field >= B2E_F_COUNT || n < 0) return fail(B2E_EINVAL, "b2e_set_rows: bad argument%s", "");
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(s->device));
  const int w = field_width(s, field), tot = n * w;
  ORDER_BEGIN(s, stream);
  B2E_LAUNCH(rows_kernel, (tot + 255) / 256, 256, 0, stream, (float*)s->fields[field], env_ids, n, w, (float*)rows, 0);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  ORDER_END(s, stream);
  return 0;
}
int b2e_get_rows(b2e_sim* s, int field, const int32_t* env_ids, int n, void* rows, void* stream) {
  if (!s || !env_ids || !rows || field < 0 || field >= B2E_F_COUNT || n < 0) return fail(B2E_EINVAL, "b2e_get_rows: bad argument%s", "");
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(s->device));
  const int w = field_width(s, field), tot = n * w;
  ORDER_BEGIN(s, stream);
  B2E_LAUNCH(rows_kernel, (tot + 255) / 256, 256, 0, stream, (float*)s->fields[field], env_ids, n, w, (float*)rows, 1);
  s->launches++;
  CUDA_TRY(cudaGetLastError());
  ORDER_END(s, stream);
  return 0;
}

int b2e_step_host(b2e_sim* s, const float* action_host, float* obs_host, float* reward_host, float* done_host,
                  int n_substeps, int mode) {
  if (!s) return fail(B2E_EINVAL, "b2e_step_host: null sim%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  const size_t na = (size_t)s->B * s->params.n_act * 4, no = (size_t)s->B * s->params.n_obs * 4, nb = (size_t)s->B * 4;
  if (mode == B2E_MODE_ACTION) {
    if (!action_host) return fail(B2E_EINVAL, "b2e_step_host: action required%s", "");
    memcpy(s->h_action, action_host, na);
    ORDER_BEGIN(s, 0);   // d_action may still be read by a launch on another stream
    CUDA_TRY(cudaMemcpyAsync(s->d_action, s->h_action, na, cudaMemcpyHostToDevice, 0));
  }
  int rc = b2e_step(s, s->d_action, obs_host ? s->d_obs : nullptr, reward_host ? s->d_reward : nullptr,
                    done_host ? s->d_done : nullptr, n_substeps, mode, 0);
  if (rc) return rc;
  if (obs_host) CUDA_TRY(cudaMemcpyAsync(s->h_obs, s->d_obs, no, cudaMemcpyDeviceToHost, 0));
  if (reward_host) CUDA_TRY(cudaMemcpyAsync(s->h_reward, s->d_reward, nb, cudaMemcpyDeviceToHost, 0));
  if (done_host) CUDA_TRY(cudaMemcpyAsync(s->h_done, s->d_done, nb, cudaMemcpyDeviceToHost, 0));
  CUDA_TRY(cudaStreamSynchronize(0));
  if (obs_host) memcpy(obs_host, s->h_obs, no);
  if (reward_host) memcpy(reward_host, s->h_reward, nb);
  if (done_host) memcpy(done_host, s->h_done, nb);
  return 0;
}

/* Host buffers that are page-locked (b2e_host_alloc): async copies go straight to/from them. */
int b2e_step_pinned(b2e_sim* s, const float* action_pinned, float* obs_pinned, float* reward_pinned, float* done_pinned,
                    int n_substeps, int mode) {
  if (!s) return fail(B2E_EINVAL, "b2e_step_pinned: null sim%s", "");
  if (mode == B2E_MODE_ACTION && !action_pinned) return fail(B2E_EINVAL, "b2e_step_pinned: action required%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  // The batch is cut into chunks that travel on two streams: the copies of one chunk overlap the kernel of
  // the next (inputs H2D -> fused step -> results D2H per chunk).
  const int na = s->params.n_act, no = s->params.n_obs;
  const int align = 2 * WPB;
  // (measured on B200: splitting a 16384-env batch costs more kernel efficiency than the overlap returns;
  //  chunks only start to pay once every chunk still fills the GPU)
  int chunks = s->B / 32768 > 1 ? (s->B / 32768 > 4 ? 4 : s->B / 32768) : 1;
  // Zero-copy results: the kernel stores observation / reward / done straight into the caller's page-locked arrays
  // (cudaMallocHost memory is device-accessible under unified addressing), so the 2.3 MB of results cross PCIe while
  // the launch's long tail is still running instead of in a copy after it (measured: 17 us instead of ~85 us between
  // the device-timed step and the end-to-end step at rollout depth 50..450).  B2ENV_ZEROCOPY=0 restores the copies.
  static const bool zero_copy = [] { const char* e = getenv("B2ENV_ZEROCOPY"); return !(e && e[0] == '0'); }();
  if (zero_copy && chunks == 1) {
    cudaStream_t st = s->pstream[0];
    ORDER_BEGIN(s, st);
    if (mode == B2E_MODE_ACTION)
      CUDA_TRY(cudaMemcpyAsync(s->d_action, action_pinned, (size_t)s->B * na * 4, cudaMemcpyHostToDevice, st));
    int rc = launch_step(s, s->d_action, obs_pinned, reward_pinned, done_pinned, n_substeps, mode, nullptr, s->B, st, 0);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
  }
  const int per = ((s->B + chunks - 1) / chunks + align - 1) / align * align;
  ORDER_BEGIN(s, s->pstream[0]);   // the chunks touch disjoint environments: ordered against earlier calls, not each other
  ORDER_BEGIN(s, s->pstream[1]);
  for (int c = 0; c * per < s->B; c++) {
    const int lo = c * per, n = (lo + per <= s->B) ? per : s->B - lo;
    cudaStream_t st = s->pstream[c & 1];
    if (mode == B2E_MODE_ACTION)
      CUDA_TRY(cudaMemcpyAsync(s->d_action + (size_t)lo * na, action_pinned + (size_t)lo * na, (size_t)n * na * 4,
                               cudaMemcpyHostToDevice, st));
    int rc = launch_step(s, s->d_action, obs_pinned ? s->d_obs : nullptr, reward_pinned ? s->d_reward : nullptr,
                         done_pinned ? s->d_done : nullptr, n_substeps, mode, nullptr, n, st, lo, false);
    if (rc) return rc;
    if (obs_pinned) CUDA_TRY(cudaMemcpyAsync(obs_pinned + (size_t)lo * no, s->d_obs + (size_t)lo * no, (size_t)n * no * 4, cudaMemcpyDeviceToHost, st));
    if (reward_pinned) CUDA_TRY(cudaMemcpyAsync(reward_pinned + lo, s->d_reward + lo, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (done_pinned) CUDA_TRY(cudaMemcpyAsync(done_pinned + lo, s->d_done + lo, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(s->pstream[0]));
  CUDA_TRY(cudaStreamSynchronize(s->pstream[1]));   // both drained: the chunks are ordered before whatever comes next
  return 0;
}

int b2e_host_alloc(void** out, size_t bytes) {
  if (!out) return fail(B2E_EINVAL, "b2e_host_alloc: null%s", "");
  CUDA_TRY(cudaMallocHost(out, bytes));
  return 0;
}
int b2e_host_free(void* p) {
  CUDA_TRY(cudaFreeHost(p));
  return 0;
}

int b2e_get(b2e_sim* s, int field, void* dst_dev, void* stream) {
  if (!s || !dst_dev || field < 0 || field >= B2E_F_COUNT) return fail(B2E_EINVAL, "b2e_get: bad argument%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  ORDER_BEGIN(s, stream);
  CUDA_TRY(cudaMemcpyAsync(dst_dev, s->fields[field], (size_t)s->B * field_width(s, field) * 4, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  ORDER_END(s, stream);
  return 0;
}
int b2e_set(b2e_sim* s, int field, const void* src_dev, void* stream) {
  if (!s || !src_dev || field < 0 || field >= B2E_F_COUNT) return fail(B2E_EINVAL, "b2e_set: bad argument%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  ORDER_BEGIN(s, stream);
  CUDA_TRY(cudaMemcpyAsync(s->fields[field], src_dev, (size_t)s->B * field_width(s, field) * 4, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  ORDER_END(s, stream);
  return 0;
}
int b2e_get_host(b2e_sim* s, int field, void* dst_host) {
  if (!s || !dst_host || field < 0 || field >= B2E_F_COUNT) return fail(B2E_EINVAL, "b2e_get_host: bad argument%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(dst_host, s->fields[field], (size_t)s->B * field_width(s, field) * 4, cudaMemcpyDeviceToHost));
  return 0;
}
int b2e_set_host(b2e_sim* s, int field, const void* src_host) {
  if (!s || !src_host || field < 0 || field >= B2E_F_COUNT) return fail(B2E_EINVAL, "b2e_set_host: bad argument%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(s->fields[field], src_host, (size_t)s->B * field_width(s, field) * 4, cudaMemcpyHostToDevice));
  return 0;
}

int b2e_debug_sched(b2e_sim* s, int32_t* envs_out, int cap, int32_t* n_main_out, int32_t* n_tail_out) {
  if (!s || !envs_out || !n_main_out || !n_tail_out) return fail(B2E_EINVAL, "b2e_debug_sched: null%s", "");
  *n_main_out = 0; *n_tail_out = 0;
  if (!s->d_sched) return 0;
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t ni = (size_t)SCHED_INTS(s->B);
  int* h = (int*)malloc(ni * sizeof(int));
  if (!h) return fail(B2E_EINVAL, "b2e_debug_sched: out of host memory%s", "");
  cudaError_t e = cudaMemcpy(h, s->d_sched, ni * sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { free(h); CUDA_TRY(e); }
  const int q = (s->sched_seq - 1) & 3, par = (s->sched_seq - 1) & 1;
  int n = 0;
  for (int b = 0; b < NBK_MAIN; b++)
    for (int i = 0; i < h[SCHED_CNT(q, b)] && n < cap; i++) envs_out[n++] = h[SCHED_LIST(par, b, s->B) + i];
  *n_main_out = n;
  const int nt = h[SCHED_CNT(q, NBK_MAIN)] < TAIL_CAP ? h[SCHED_CNT(q, NBK_MAIN)] : TAIL_CAP;
  for (int i = 0; i < nt && n < cap; i++) envs_out[n++] = h[SCHED_TAIL(par, s->B) + i];
  *n_tail_out = n - *n_main_out;
  free(h);
  return 0;
}

int b2e_timer_start(b2e_sim* s, void* stream) {
  if (!s) return fail(B2E_EINVAL, "b2e_timer_start: null%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaEventRecord(s->ev0, (cudaStream_t)stream));
  return 0;
}
int b2e_timer_stop(b2e_sim* s, void* stream, float* ms_out) {
  if (!s || !ms_out) return fail(B2E_EINVAL, "b2e_timer_stop: null%s", "");
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaEventRecord(s->ev1, (cudaStream_t)stream));
  CUDA_TRY(cudaEventSynchronize(s->ev1));
  CUDA_TRY(cudaEventElapsedTime(ms_out, s->ev0, s->ev1));
  return 0;
}

}  // extern "C"

#ifdef PROFILE_WARM
extern "C" int b2e_debug_lane_sites(unsigned long long* out32, int reset) {   // probe build only (not part of include/b2env.h)
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out32, g_lane_site, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_lane_site, z, sizeof(z)); }
  return 0;
}
#endif
