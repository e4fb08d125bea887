/*
 * b2env.h — C-ABI of the B200-native batched step + reward pipeline.
 *
 * This is the drop-in boundary for the ONE hot path of hsp-iit/pybullet-robot-envs:
 * what the reference reaches through `import pybullet as p` during env.step()/reset():
 *
 *   p.stepSimulation                 panda_push_gym_env.py:236, panda_reach_gym_env.py:220
 *   p.setJointMotorControl2          panda_env.py:305-310 (joint mode targets)
 *   p.setJointMotorControlArray      panda_env.py:276-282 (IK mode targets)
 *   p.calculateInverseKinematics     panda_env.py:269-272
 *   p.getLinkState/getJointStates    panda_env.py:147, :187
 *   p.getBasePositionAndOrientation  world_env.py:114
 *   p.getEulerFromQuaternion / getQuaternionFromEuler / invertTransform /
 *   p.multiplyTransforms             panda_push_gym_env.py:168-176
 *   p.resetSimulation/loadURDF/resetJointState (the reset path)
 *                                    panda_push_gym_env.py:117-148, panda_env.py:51-91,
 *                                    world_env.py:61-84
 *
 * Plain C, no torch types: device pointers are `void*`/typed pointers into CUDA
 * global memory, the stream is a `cudaStream_t` passed as `void*`.
 * Every function returns 0 on success or a negative B2E_E* code; the message is
 * available from b2e_last_error() (thread-local).  Nothing throws across the ABI.
 *
 * The same structs are consumed by the CPU oracle (oracle/b2oracle.c), which is
 * test infrastructure and never linked into this library.
 */
#ifndef B2ENV_H
#define B2ENV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2E_MAX_LINKS 40   /* iCub: 38 joints */
#define B2E_MAX_DOF 32
#define B2E_MAX_SPHERES 16
#define B2E_MAX_OBS 40
#define B2E_MAX_CONTACTS 12 /* contact points kept per env per step            */
#define B2E_MAX_BOXES 4     /* box collision proxies of robot links (finger pads) */
#define B2E_MAX_SELF_PAIRS 32 /* sphere pairs tested for robot self-collision    */
#define B2E_MAX_SBOXES 6    /* static world boxes (table top, legs)             */
#define B2E_MAX_CAPS 4      /* capsule collision proxies of robot links          */
#define B2E_MAX_LIMROWS 3   /* joint-limit rows kept per env per step          */
#define B2E_CACHE_SLOTS 16  /* warm-start cache slots (key + 3 impulses each)  */

/* joint types (PyBullet numbering: JOINT_REVOLUTE 0, JOINT_PRISMATIC 1, JOINT_FIXED 4) */
#define B2E_JOINT_REVOLUTE 0
#define B2E_JOINT_PRISMATIC 1
#define B2E_JOINT_FIXED 4

/* error codes */
#define B2E_OK 0
#define B2E_EINVAL (-1)
#define B2E_ECUDA (-2)
#define B2E_ENOMEM (-3)
#define B2E_EUNSUPPORTED (-4)

/* Articulated model descriptor (what p.loadURDF builds into a btMultiBody).
 * Link i is the child of joint i; joint/link index == PyBullet joint index
 * (URDF order, panda_env.py:60-79 relies on it).  parent == -1 is the base.   */
typedef struct b2e_model {
  int32_t n_links;                 /* joints (= non-base links)                 */
  int32_t n_dof;                   /* movable joints                            */
  int32_t ee_link;                 /* end-effector link (panda_env.py:40 -> 11) */
  int32_t n_spheres;               /* collision proxies                         */
  int32_t parent[B2E_MAX_LINKS];
  int32_t jtype[B2E_MAX_LINKS];
  int32_t dof[B2E_MAX_LINKS];      /* dof index of joint i, or -1               */
  float jpos[B2E_MAX_LINKS][3];    /* joint origin in parent link frame         */
  float jrot[B2E_MAX_LINKS][9];    /* rotation parent<-joint frame, row-major   */
  float axis[B2E_MAX_LINKS][3];    /* joint axis in joint (=child link) frame   */
  float mass[B2E_MAX_LINKS];
  float com[B2E_MAX_LINKS][3];     /* COM in link frame                         */
  float inertia[B2E_MAX_LINKS][9]; /* about COM, link-frame axes, row-major     */
  float lower[B2E_MAX_DOF];
  float upper[B2E_MAX_DOF];
  float limit_margin[B2E_MAX_DOF]; /* limit rows are instantiated inside this   */
  float max_force[B2E_MAX_DOF];    /* motor max force (N or N*m)                */
  float max_vel[B2E_MAX_DOF];      /* motor velocity clamp, <=0: none           */
  float joint_damping[B2E_MAX_DOF];
  float home[B2E_MAX_DOF];         /* panda_env.py:19-23                        */
  float base_pos[3];               /* fixed base pose in world                  */
  float base_rot[9];
  int32_t sph_link[B2E_MAX_SPHERES]; /* owning link                             */
  float sph_c[B2E_MAX_SPHERES][3];   /* centre in link frame                    */
  float sph_r[B2E_MAX_SPHERES];
  float sph_mu[B2E_MAX_SPHERES];     /* lateral friction of the owning link     */
  float sph_erp[B2E_MAX_SPHERES];    /* <0: use global erp                      */
  float sph_cfm[B2E_MAX_SPHERES];
  /* box proxies (finger pads): collided with the object and the static world by box-box SAT + face clipping
   * (include/b2env_narrowphase.h), the narrowphase Bullet runs for box shapes                                  */
  int32_t n_boxes;
  int32_t box_link[B2E_MAX_BOXES];
  float box_c[B2E_MAX_BOXES][3];     /* centre in link frame                      */
  float box_h[B2E_MAX_BOXES][3];     /* half extents along the link axes          */
  float box_mu[B2E_MAX_BOXES];
  float box_erp[B2E_MAX_BOXES];      /* <0: use global erp                        */
  float box_cfm[B2E_MAX_BOXES];
  /* robot self-collision (URDF_USE_SELF_COLLISION, panda_env.py:53): candidate pairs of sphere proxies on links that
   * are neither the same nor parent / child (Bullet's default filter), built on the host                       */
  int32_t n_self_pairs;
  int32_t self_a[B2E_MAX_SELF_PAIRS];
  int32_t self_b[B2E_MAX_SELF_PAIRS];
  /* sphere flags: bit 0 = self-collision partner only (not tested against the object or the static world: the link's
   * world contacts come from another proxy, or the link cannot reach the world)                                  */
  int32_t sph_flags[B2E_MAX_SPHERES];
  /* capsule proxies (segment p0-p1 in the link frame + radius): collided with the object and the static boxes by GJK /
   * EPA on the segment core (include/b2env_narrowphase.h: b2n_convex_contact), with the ground plane in closed form.
   * This is the convex-convex narrowphase Bullet runs for hull / capsule shapes (btGjkPairDetector + EPA)          */
  int32_t n_caps;
  int32_t cap_link[B2E_MAX_CAPS];
  float cap_p0[B2E_MAX_CAPS][3];
  float cap_p1[B2E_MAX_CAPS][3];
  float cap_r[B2E_MAX_CAPS];
  float cap_mu[B2E_MAX_CAPS];
} b2e_model;

#define B2E_TASK_REACH 0
#define B2E_TASK_PUSH 1
#define B2E_TASK_GRASP 2 /* PandaGrasp-v0 (BASELINE.json config 5; no reference env: built from the grasp
                            primitives panda_env.py:195-225 and helloworld_panda.py:94-148)          */

/* World + solver + task constants (what the task env and WorldEnv set up).     */
typedef struct b2e_params {
  float dt;                /* 1/240, panda_push_gym_env.py:39                   */
  float gravity[3];        /* (0,0,-9.8), :126                                   */
  int32_t solver_iters;    /* 150, :122                                          */
  float residual_tol;      /* squared velocity residual for early exit           */
  float erp;               /* contact / limit error reduction                   */
  float slop;              /* linear slop                                        */
  float warmstart;         /* warm-start factor for contact impulses            */
  float contact_margin;    /* contacts generated when distance < margin         */
  float table_min[3];      /* table-top slab (static box)                        */
  float table_max[3];
  float table_mu;
  float plane_mu;          /* ground plane z = 0                                 */
  float cube_half;         /* half extent of cube_small                          */
  float cube_mass;
  float cube_inertia;      /* isotropic box inertia m*(2a)^2/6                   */
  float cube_mu;
  float damp_lin_k1, damp_lin_k2, damp_ang_k1, damp_ang_k2; /* free-body damping */
  int32_t task;            /* B2E_TASK_*                                         */
  int32_t n_act;           /* action width (7 joint mode; 3/6 IK mode)           */
  int32_t n_ctrl;          /* numControlledJoints (joint mode)                   */
  int32_t use_ik;
  int32_t ik_orientation;  /* control_orientation                                */
  int32_t ik_iters;        /* 100, panda_env.py:270                              */
  float ik_residual;       /* 1e-3, panda_env.py:271                             */
  float ik_damping;        /* DLS damping                                        */
  float act_scale;         /* 0.05 joint mode (:225)                             */
  float act_scale_pos;     /* 0.005 (:199)                                       */
  float act_scale_rot;     /* 0.01  (:203)                                       */
  float kp_ctrl;           /* 0.5, panda_env.py:308                              */
  float kp_hold;           /* 0.2, panda_env.py:76                               */
  float dist_min;          /* success radius: 0.1 push (:52), 0.03 reach         */
  int32_t max_steps;
  int32_t n_obs;
  float obs_low[B2E_MAX_OBS];
  float obs_high[B2E_MAX_OBS];
  float vel_mean[3];       /* panda_env.py:175                                   */
  float vel_std[3];        /* panda_env.py:174                                   */
  float ws_lim[3][2];      /* robot workspace (IK clamps)                        */
  float eu_lim[3][2];
  float home_hand_pose[6]; /* panda_env.py:85-88                                 */
  float kp_grip;           /* GRASP: finger position gain (setJointMotorControl2 default 0.1, panda_env.py:218-224) */
  float grasp_lift;        /* GRASP: success when the object is this far above its rest height    */
  float grasp_rest_z;      /* GRASP: object rest height (table top + half extent)                 */
  int32_t goal_env;        /* GoalEnv semantics (panda_push_gym_goal_env.py:89-122): sparse reward
                              -(d > dist_min), done = counter > max_steps or success, no latch */
  /* ---- iCub task envs (icub_envs/icub_env.py, icub_push_gym_env.py, icub_reach_gym_env.py); all zero = Panda ---- */
  int32_t n_obs_joints;    /* joints listed in the robot observation (_joints_to_control, icub_env.py:242-247);
                              0: every dof (panda_env.py:187-191)                                          */
  int32_t obs_dof[16];     /* their dof indices, in joint-index order                                      */
  int32_t ctrl_dof[16];    /* joint mode: action k drives dof ctrl_dof[k] (icub_env.py:347-361); only read
                              when n_obs_joints > 0 (Panda: dof k)                                         */
  uint32_t ctrl_mask;      /* bit d: dof d is controlled.  IK mode: the other joints are blocked, their target
                              is the rest pose (icub_env.py:314-317).  Only read when n_obs_joints > 0      */
  float ik_link_offset[3]; /* COM -> link frame of the hand: the IK target is the commanded pose composed with
                              this translation (icub_env.py:251-257, :303-305)                             */
  int32_t reward_kind;     /* B2E_REWARD_*                                                                 */
  int32_t max_contacts;    /* contact points kept per env per step; <= 0: B2E_MAX_CONTACTS                 */
  /* ---- static world boxes (world_env.py:62-66: table/table.urdf = top slab + four legs; [0] is the top slab and
   *      equals table_min / table_max); the ground plane z = 0 is implicit.  n_sboxes = 0: top slab only          */
  int32_t n_sboxes;
  float sbox_c[B2E_MAX_SBOXES][3];   /* centre (world, axis aligned)             */
  float sbox_h[B2E_MAX_SBOXES][3];   /* half extents                             */
  float sbox_mu[B2E_MAX_SBOXES];
  /* ---- Cartesian control with a velocity cap: robot.apply_action(action, max_vel) with max_vel != -1 (panda_env.py:285-291,
   *      icub_env.py:331-337) drives the n_ctrl controlled joints through setJointMotorControl2(maxVelocity = max_vel) with
   *      PyBullet's DEFAULT position gain instead of setJointMotorControlArray(positionGains = 0.2).  <= 0: off          */
  float ik_max_vel;
  float kp_ik_max_vel;     /* PyBullet default positionGain 0.1 [EXT-recalled]                                       */
} b2e_params;

#define B2E_REWARD_PANDA 0       /* panda_push_gym_env.py:318-331 / panda_reach_gym_env.py:303-313 (bonus replaces) */
#define B2E_REWARD_ICUB_REACH 1  /* icub_reach_gym_env.py:319-330: -d, bonus 1000 + (100 - 80 d) ADDED             */
#define B2E_REWARD_ICUB_PUSH0 2  /* icub_push_gym_env.py:353-356: -d1 - d2, +1000 on success                       */
#define B2E_REWARD_ICUB_PUSH1 3  /* icub_push_gym_env.py:358-371: shaping normalised by the distances at reset      */

/* state fields for b2e_get / b2e_set (all env-major, [B][width])               */
enum b2e_field {
  B2E_F_Q = 0,        /* float [B][n_dof]   joint positions                     */
  B2E_F_QD = 1,       /* float [B][n_dof]   joint velocities                    */
  B2E_F_OBJ_POSE = 2, /* float [B][7]       xyz + quaternion xyzw               */
  B2E_F_OBJ_VEL = 3,  /* float [B][6]       linear, angular (world)             */
  B2E_F_TARGET = 4,   /* float [B][3]                                           */
  B2E_F_MTARGET = 5,  /* float [B][n_dof]   motor position targets              */
  B2E_F_COUNTERS = 6, /* int32 [B][2]       env_step_counter, terminated        */
  B2E_F_CACHE_KEY = 7,/* int32 [B][16]      contact feature keys (-1 empty)     */
  B2E_F_CACHE_LAM = 8,/* float [B][16][3]   normal, friction1, friction2 impulse*/
  B2E_F_HAND_POSE = 9,/* float [B][6]       IK-mode commanded hand pose         */
  B2E_F_STATUS = 10,  /* int32 [B][4]       flags, pgs iters, n_contacts, n_rows*/
  B2E_F_RAW_OBS = 11, /* float [B][n_obs]   unscaled observation of last step   */
  B2E_F_CONTACTS = 12,/* float [B][12][8]   last step: key, dist, n(3), lam(3)  */
  B2E_F_SHAPING = 13, /* float [B][2]       _init_dist_hand_obj, _max_dist_obj_tg captured at reset
                                            (icub_push_gym_env.py:126-127)       */
  B2E_F_COUNT = 14
};

/* Contact feature keys (B2E_F_CACHE_KEY, B2E_F_CONTACTS): stable ids used by the warm-start cache.
 *   0..7    cube vertex v vs table top (cube fully over the table: vertex-face manifold)
 *   8..15   cube vertex v vs ground plane
 *   16..31  robot sphere s vs cube
 *   32..47  robot sphere s vs table top
 *   64..95  robot box b vertex v vs table top (64 + 8 b + v);  96..127 the same vs the ground plane
 *   128..255 robot sphere s vs static box k other than through the table top (128 + 8 s + k)
 *   256..271 robot sphere s vs ground plane
 *   512..543 robot self-collision, sphere pair p
 *   576..579 robot capsule c vs cube (GJK / EPA);  592..623 capsule c vs static box k (592 + 8 c + k);
 *   624..631 capsule c end e vs ground plane (624 + 2 c + e)
 *   4096 + 1024 k + id   cube vs static box k, general box-box (rim, legs; id: include/b2env_narrowphase.h)
 *   12288 + 1024 b + id  robot box b vs cube, box-box
 *   16384 + 1024 (8 b + k) + id  robot box b vs static box k, general box-box                                   */
#define B2E_KEY_CUBE_TABLE 0
#define B2E_KEY_CUBE_PLANE 8
#define B2E_KEY_SPHERE_CUBE 16
#define B2E_KEY_SPHERE_TABLE 32
#define B2E_KEY_BOXV_TABLE 64
#define B2E_KEY_BOXV_PLANE 96
#define B2E_KEY_SPHERE_SBOX 128
#define B2E_KEY_SPHERE_PLANE 256
#define B2E_KEY_SELF 512
#define B2E_KEY_CAP_CUBE 576
#define B2E_KEY_CAP_SBOX 592
#define B2E_KEY_CAP_PLANE 624
#define B2E_KEY_CUBE_SBOX 4096
#define B2E_KEY_BOX_CUBE 12288
#define B2E_KEY_BOX_SBOX 16384

/* status flag bits */
#define B2E_ST_NAN 1
#define B2E_ST_CONTACT_OVERFLOW 2
#define B2E_ST_LIMIT_OVERFLOW 4

/* step modes */
#define B2E_MODE_ACTION 0 /* targets from the action (apply_action)             */
#define B2E_MODE_HOLD 1   /* keep motor targets, hold gains (settle steps of reset) */
#define B2E_MODE_IK_POSE 3 /* Cartesian mode: IK of the stored hand pose -> motor targets, no action
                              increment, no task bookkeeping (robot.reset, panda_env.py:83-91) */
#define B2E_MODE_OBSERVE 4 /* n_substeps = 0: observation / reward / done of the current state, nothing is
                              written back (get_extended_observation, _compute_reward); with
                              B2E_MODE_HOLD the success latch `terminated` is stored like _termination() */
#define B2E_MODE_TARGETS 2 /* keep motor targets already written to B2E_F_MTARGET by
                              robot.apply_action, control gains; no task bookkeeping */

typedef struct b2e_sim b2e_sim;

/* Create a simulation of num_envs environments on CUDA device `device`.
 * Descriptors are copied.  Replaces p.connect + loadURDF (panda_push_gym_env.py:56-70). */
int b2e_create(const b2e_model* model, const b2e_params* params, int num_envs, int device,
               b2e_sim** out);
void b2e_destroy(b2e_sim* sim);

/* Update the constants (e.g. kp switch between settle and control).            */
int b2e_set_params(b2e_sim* sim, const b2e_params* params);

/* Options. B2E_OPT_RECORD_CONTACTS: also write the per-step contact list
 * (B2E_F_CONTACTS; what p.getContactPoints reports, panda_env.py:315,327-328).  */
#define B2E_OPT_RECORD_CONTACTS 0
int b2e_set_option(b2e_sim* sim, int option, int value);

/* Masked reset: for every env with env_mask[e] != 0 (NULL = all) load the home
 * joint state (resetJointState, panda_env.py:72), zero velocities, motor targets
 * = home, object pose = obj_init_pose[e] (world_env.py:81-84), target, counters,
 * empty contact cache.  All pointers are device pointers.  No settle steps —
 * call b2e_step(mode=HOLD) for those (panda_push_gym_env.py:132-148).           */
int b2e_reset(b2e_sim* sim, const uint8_t* env_mask, const float* obj_init_pose /*[B][7]*/,
              const float* target /*[B][3]*/, void* stream);

/* One env.step(): action -> motor targets -> n_substeps x (dynamics, collision,
 * PGS, integrate) -> observation, reward, done
 * (panda_push_gym_env.py:244-255).  action [B][n_act], obs [B][n_obs] (scaled
 * with scale_gym_data, utils.py:78-91), reward [B], done [B] (float 0/1).
 * obs/reward/done may be NULL (settle steps).  All device pointers.
 * n_substeps == 0 with mode HOLD only evaluates observation / reward / done on
 * the current state (get_extended_observation, _termination, _compute_reward).
 * Launches over the whole batch of a Panda-class model reorder their blocks by the
 * cost of the previous solve (slow blocks first); results do not depend
 * on it.  Environment variable B2ENV_SCHED=0 (read at b2e_create) switches it off.  */
int b2e_step(b2e_sim* sim, const float* action, float* obs, float* reward, float* done,
             int n_substeps, int mode, void* stream);

/* The same step for a SUBSET of environments (env_ids: device int32 [n_ids], distinct): per-env
 * resets settle only the envs being reset (panda_push_gym_env.py:105-148 applied to some envs of
 * the batch).  action/obs/reward/done are the full-batch buffers, indexed by env id.             */
int b2e_step_subset(b2e_sim* sim, const int32_t* env_ids, int n_ids, const float* action, float* obs,
                    float* reward, float* done, int n_substeps, int mode, void* stream);
/* Scatter / gather rows of a state field: rows is a device buffer [n][width].                    */
int b2e_set_rows(b2e_sim* sim, int field, const int32_t* env_ids, int n, const void* rows, void* stream);
int b2e_get_rows(b2e_sim* sim, int field, const int32_t* env_ids, int n, void* rows, void* stream);

/* Stream contract: every entry point that takes `stream` enqueues its work there; the host-buffer entry
 * points below use the library's own streams.  Calls on one b2e_sim are ORDERED whatever streams they use:
 * each launch records an event that the next entry point's stream waits on, so a caller on a non-blocking
 * (e.g. torch side) stream may freely mix b2e_step, b2e_step_pinned, b2e_set_rows, ...  A b2e_sim is still
 * not thread-safe: drive it from one host thread at a time.                      */

/* Host-buffer convenience used by the single-env compatibility path and the
 * end-to-end benchmark: copies action H2D, steps, copies results D2H, syncs.    */
int b2e_step_host(b2e_sim* sim, const float* action_host, float* obs_host, float* reward_host,
                  float* done_host, int n_substeps, int mode);

/* Same, for PAGE-LOCKED host buffers (from b2e_host_alloc): the async copies go straight
 * to / from the caller's memory, no staging copy.                               */
int b2e_step_pinned(b2e_sim* sim, const float* action_pinned, float* obs_pinned, float* reward_pinned,
                    float* done_pinned, int n_substeps, int mode);
int b2e_host_alloc(void** out, size_t bytes);
int b2e_host_free(void* p);

/* Copy a state field to / from a DEVICE buffer of the field's full size.        */
int b2e_get(b2e_sim* sim, int field, void* dst_dev, void* stream);
int b2e_set(b2e_sim* sim, int field, const void* src_dev, void* stream);
/* Same with HOST buffers (synchronous).                                         */
int b2e_get_host(b2e_sim* sim, int field, void* dst_host);
int b2e_set_host(b2e_sim* sim, int field, const void* src_host);

/* Diagnostics (tests): the scheduling lists written by the last scheduled full-batch step — the environments of the
 * main launch's cost classes (lightest class first), then those of the tail launch.  Every environment appears exactly
 * once.  n_main = n_tail = 0 when the simulation is not scheduled (small batches, the tree kernel).               */
int b2e_debug_sched(b2e_sim* sim, int32_t* envs_out, int cap, int32_t* n_main_out, int32_t* n_tail_out);

/* Width (elements per env) and element size of a field.                        */
int b2e_field_width(const b2e_sim* sim, int field);
int b2e_field_elem_size(int field);

int b2e_num_envs(const b2e_sim* sim);
/* Number of kernels this library launched since creation (bench gpu_launches). */
int64_t b2e_launch_count(const b2e_sim* sim);
/* Average device time (ms) of the step kernel between two marks: timed with
 * CUDA events recorded on the launch stream.                                    */
int b2e_timer_start(b2e_sim* sim, void* stream);
int b2e_timer_stop(b2e_sim* sim, void* stream, float* ms_out);

const char* b2e_last_error(void);
const char* b2e_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B2ENV_H */
