/*
 * trl_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A from-scratch, plain-C restatement of the reference's (UBCMOCCA/TerrainRLSim)
 * double-precision algorithms for the batched control-and-observation step.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (terrainrlsim_b200/, include/) never does.
 *
 * PARITY PINNED TO THE REFERENCE ITSELF: the reference ships no golden vectors, but its own
 * translation units compile unmodified against stand-in Eigen / Bullet headers into
 * oracle/_ref (oracle/Makefile `ref`); tests/test_ref_pin.py compares this oracle with it at
 * 1e-13 (most results to the last bit) and tests/golden/ freezes the reference outputs.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference tree). Third-party arithmetic restated from its published algorithm:
 * Eigen 3.3 LDLT (pivoted, in-place lower) and Quaternion::slerp -- the reference
 * does not pin an Eigen version.
 */
#ifndef TRL_ORACLE_H
#define TRL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define TRLO_JOINT_COLS 19 /* cKinTree::eJointDescMax, anim/KinTree.h:24-46 */
#define TRLO_BODY_COLS 17  /* cKinTree::eBodyParamMax, anim/KinTree.h:57-77 */
#define TRLO_PD_COLS 18    /* cPDController::eParamMax, sim/PDController.h:12-32 */
#define TRLO_MAX_JOINTS 64
#define TRLO_MAX_DOF 128

/* joint types, anim/KinTree.h:13-22 */
enum { TRLO_JOINT_REVOLUTE = 0, TRLO_JOINT_PLANAR = 1, TRLO_JOINT_PRISMATIC = 2,
       TRLO_JOINT_FIXED = 3, TRLO_JOINT_SPHERICAL = 4, TRLO_JOINT_NONE = 5 };

typedef struct trlo_model {
    int num_joints;            /* N */
    int num_dof;               /* n = cKinTree::GetNumDof */
    const double* joint_mat;   /* [N][19] row-major */
    const double* body_defs;   /* [N][17] */
    const double* pd_params;   /* [N][18] */
    double gravity[3];
    int num_frames;            /* 0 = no motion */
    const double* frames;      /* [F][1+n], col 0 = CUMULATIVE start time (post cMotion::PostProcessFrames) */
    int loop;
} trlo_model;

/* ground description: the caller supplies origin/scale/bounds as doubles */
enum { TRLO_GROUND_NONE = 0, TRLO_GROUND_VAR2D = 1, TRLO_GROUND_VAR3D = 2, TRLO_GROUND_PLANE = 3 };
typedef struct trlo_ground {
    int kind;
    int num_pieces;            /* 2 segments (2-D) or 4 slabs (3-D) */
    const float* data[4];      /* heights as stored (already multiplied by world_scale as float) */
    int res[4][2];             /* grid width (x), grid length (z) */
    double origin[4][3];       /* cSimObj::GetPos() of the piece */
    double scale[4][3];        /* GetScaling(): bullet local scaling / world_scale */
    double bounds[4][4];       /* min_x, max_x, min_z, max_z used by ContainsPos */
} trlo_ground;

/* observation configuration */
enum { TRLO_OBS_TERRAIN_RL = 0,        /* cTerrainRLCharController base features */
       TRLO_OBS_TERRAIN_RL_HOPPER = 1, /* + cMonopedHopperController overrides */
       TRLO_OBS_CT = 2,                /* cCtController 2-D features */
       TRLO_OBS_CT_3D = 3 };           /* cCtController 3-D features with ENABLE_MAX_STATE */
enum { TRLO_SAMPLER_LINE = 0, TRLO_SAMPLER_GRID = 1 };
typedef struct trlo_obs_cfg {
    int obs_kind;
    int sampler;        /* line: G samples on x in [view_min, view_max]; grid: res x res square */
    int num_samples;    /* G (grid: res*res) */
    int grid_res;
    double view_min;    /* line: gViewMin=-0.5; grid: min view dist */
    double view_max;    /* line: view dist (10); grid: max view dist */
    int num_phase_bins; /* <0: no phase block; >=0: 1 + bins entries (cCtPhaseController) */
} trlo_obs_cfg;

/* ---- model queries ---- */
int trlo_param_offset(const trlo_model* m, int j);
int trlo_param_size(const trlo_model* m, int j);
int trlo_joint_type(const trlo_model* m, int j);
int trlo_parent(const trlo_model* m, int j);
int trlo_valid_body(const trlo_model* m, int j);
int trlo_pd_valid(const trlo_model* m, int j);

/* ---- kinematics (anim/KinTree.cpp) ---- */
void trlo_child_parent_trans(const trlo_model* m, const double* pose, int j, double out[16]);
void trlo_joint_world_trans(const trlo_model* m, const double* pose, int j, double out[16]);
void trlo_body_part_pos(const trlo_model* m, const double* pose, int j, double out[3]);
void trlo_vel_to_pose_diff(const trlo_model* m, const double* pose, const double* vel, double* out);
void trlo_post_process_pose(const trlo_model* m, double* pose);
void trlo_calc_vel(const trlo_model* m, const double* pose0, const double* pose1, double dt, double* out_vel);
void trlo_lerp_poses(const trlo_model* m, const double* pose0, const double* pose1, double lerp, double* out);
double trlo_calc_heading(const double* pose);
void trlo_build_origin_trans(const double* pose, double out[16]);

/* ---- rigid body dynamics (sim/RBDModel.cpp, sim/RBDUtil.cpp, sim/SpAlg.cpp) ---- */
/* out_mass [n][n] row-major, out_bias [n]; either may be NULL */
void trlo_rbd_update(const trlo_model* m, const double* pose, const double* vel,
                     double* out_mass, double* out_bias);
/* full inverse dynamics with given acc (cRBDUtil::SolveInvDyna); used by identity tests */
void trlo_solve_inv_dyna(const trlo_model* m, const double* pose, const double* vel,
                         const double* acc, double* out_tau);
void trlo_inertia_spatial(const trlo_model* m, int j, double out[36]);
/* RBD extras: cRBDUtil::BuildJacobian / BuildCOMJacobian / CalcGravityForce
 * (sim/RBDUtil.cpp:256-275,294-345,937-982); J, Jcom row-major [6][n]; any output may be NULL */
void trlo_rbd_extras(const trlo_model* m, const double* pose, double* out_J, double* out_Jcom, double* out_gravity_force);
/* Torque post-processing: cSimCharacter::ApplyControlForces + cJoint::ApplyTau with
 * ClampTotalTorque / ClampTotalForce (sim/SimCharacter.cpp:589-603, sim/Joint.cpp:162-221,300-318,536-562).
 * out_tau [n] clamped generalized torques, out_body_wrench [N][6] world (torque, force) per body; either may be NULL */
void trlo_apply_control_forces(const trlo_model* m, const double* pose, const double* tau, double* out_tau,
                               double* out_body_wrench);
/* cRBDUtil::BuildEndEffectorJacobian(model, joint_id), sim/RBDUtil.cpp:181-206; row-major [6][n] */
void trlo_end_eff_jacobian(const trlo_model* m, const double* pose, int joint_id, double* out_J);

/* Eigen 3.3 LDLT restated: A [n][n] row-major symmetric (only lower read), solves A x = b.
 * returns number of zero pivots. */
int trlo_ldlt_solve(int n, const double* A, const double* b, double* x);

/* ---- stable PD (sim/ImpPDController.cpp:137-196) ---- */
/* kp_joint/kd_joint: [N] per-joint gains or NULL (use pd_params); joint_active [N] or NULL (all active);
 * tar_pose/tar_vel [n]; inout_tau [n] accumulated (+=); out_mass/out_bias optional. returns #zero pivots.
 * Targets of joints whose pd_params row has UseWorldCoord != 0 are WORLD-coordinate targets and are converted with
 * the current joint rotations like cPDController::GetTargetTheta does (sim/PDController.cpp:223-273). */
int trlo_imp_pd_control_force(const trlo_model* m, double dt, const double* pose, const double* vel,
                              const double* tar_pose, const double* tar_vel,
                              const unsigned char* joint_active,
                              const double* kp_joint, const double* kd_joint,
                              double* inout_tau, double* out_mass, double* out_bias);

/* explicit PD: cExpPDController::UpdateControlForce -> cPDController::CalcJointTau* (sim/ExpPDController.cpp:59-66,
 * 148-159; sim/PDController.cpp:134-149,302-428). Same argument meaning as the stable-PD call. */
void trlo_exp_pd_control_force(const trlo_model* m, double dt, const double* pose, const double* vel,
                               const double* tar_pose, const double* tar_vel,
                               const unsigned char* joint_active,
                               const double* kp_joint, const double* kd_joint, double* inout_tau);

/* ---- motion / kinematic character (anim/Motion.cpp, anim/KinCharacter.cpp) ---- */
void trlo_motion_index_phase(const trlo_model* m, double time, int* out_idx, double* out_phase);
void trlo_motion_calc_frame(const trlo_model* m, double time, double* out_pose);
/* cKinCharacter::Pose(t): pose at t (with cycle root delta, origin rot/pos) and finite-diff vel */
void trlo_kin_char_pose(const trlo_model* m, double time, const double origin_pos[3],
                        const double origin_rot[4] /* w,x,y,z */, double* out_pose, double* out_vel);

/* ---- ground (sim/GroundVar2D.cpp, sim/GroundVar3D.cpp) ---- */
/* out_cell: piece, i, j (piece=-1 if no piece contains pos) */
double trlo_sample_height(const trlo_ground* g, const double pos[3], int* out_valid, int out_cell[3]);
void trlo_sample_height_batch(const trlo_ground* g, int count, const double* pos, double* out_h, int* out_valid,
                              int* out_cell);

/* ---- observation (sim/TerrainRLCharController.cpp, sim/CtController.cpp, ...) ---- */
int trlo_poli_state_size(const trlo_model* m, const trlo_obs_cfg* cfg);
void trlo_ground_sample_trans(const trlo_ground* g, const double* pose, double out[16]);
void trlo_parse_ground(const trlo_ground* g, const trlo_obs_cfg* cfg, const double* pose, int flip,
                       double* out_samples, int* out_cells /* [G][3] or NULL */);
/* vel [n] is used only by the HOPPER override (root omega z); body_pos/body_vel [N][3];
 * body_quat [N][4] (w,x,y,z) and body_angvel [N][3] only for CT_3D; ground_vel [3] or NULL */
void trlo_build_poli_state(const trlo_model* m, const trlo_ground* g, const trlo_obs_cfg* cfg,
                           const double* pose, const double* vel, const double* body_pos, const double* body_vel,
                           const double* body_quat, const double* body_angvel, const double* ground_vel,
                           double phase, int flip, double* out_state);

/* ---- imitation reward (scenarios/ScenarioExpImitate.cpp:13-159); pose0/vel0 = simulated character, pose1/vel1 =
 * kinematic reference, body_vel [N][3] = body linear velocities (for cSimCharacter::CalcCOMVel); end-effector world
 * positions of the simulated character are the FK of pose0. out_terms [5] (pose, vel, end-eff, root, com errors). ---- */
/* cRBDUtil::CalcCoM(joint_mat, body_defs, pose, vel, com, com_vel), sim/RBDUtil.cpp:557-598 */
void trlo_calc_com(const trlo_model* m, const double* pose, const double* vel, double com[3], double com_vel[3]);
double trlo_imitate_reward(const trlo_model* m, const trlo_ground* g, const double* pose0, const double* vel0,
                           const double* pose1, const double* vel1, const double* body_vel, int fallen,
                           double* out_terms);

/* ---- terrain generation and scrolling (oracle/trl_oracle_terrain.c) ---- */
/* cRand (util/Rand.cpp) on std::default_random_engine = minstd_rand0 */
typedef struct trlo_rand { unsigned long long x; } trlo_rand;
void trlo_rand_seed(trlo_rand* r, unsigned long seed);
double trlo_rand_double(trlo_rand* r);
double trlo_rand_double_range(trlo_rand* r, double mn, double mx);
int trlo_rand_int(trlo_rand* r);
int trlo_rand_int_range(trlo_rand* r, int mn, int mx);
int trlo_rand_flip_coin(trlo_rand* r, double p);
/* cTerrainGen2D: type = cGround::eType (sim/Ground.h:26-55: 1 flat, 2 gaps, 3 steps, 4 walls, 5 bumps, 6 mixed,
 * 7 narrow gaps, 8..13 slopes variants, 14 cliffs); params [trlo_terrain2d_num_params()] or NULL for the defaults;
 * appends to data[*inout_count ...]; returns the width added */
int trlo_terrain2d_num_params(void);
void trlo_terrain2d_default_params(double* out);
double trlo_terrain_gen_2d(int type, double width, const double* params, trlo_rand* rand, float* data,
                           int* inout_count, int cap);
/* one terrain piece (2-D segment / 3-D slab) with its Bullet-side placement */
#define TRLO_TERRAIN_PIECE_CAP 49152
typedef struct trlo_terrain_piece {
    float data[TRLO_TERRAIN_PIECE_CAP];
    int count;          /* floats in data (2-D: 3 rows of res[0]) */
    int res[2];
    double origin[3];   /* cSimObj::GetPos() */
    double scale[3];    /* GetScaling() */
    double aabb[4];     /* rigid-body AABB min_x, max_x, min_z, max_z (/ world scale) */
    double min[3], max[3]; /* 3-D: tSlab::mMin / mMax */
} trlo_terrain_piece;
typedef struct trlo_ground_var2d {
    int type;
    double params[40];
    double ground_width, world_scale;
    trlo_rand rand;
    int flip_seg;
    trlo_terrain_piece piece[2];
} trlo_ground_var2d;
void trlo_ground2d_init(trlo_ground_var2d* g, int type, const double* params, double ground_width, double world_scale,
                        unsigned long seed, double bound_min_x, double bound_max_x);
void trlo_ground2d_update(trlo_ground_var2d* g, double bound_min_x, double bound_max_x);
void trlo_ground2d_view(const trlo_ground_var2d* g, trlo_ground* out);
typedef struct trlo_ground_var3d {
    double ground_width, spacing_x, spacing_z, world_scale;
    int slab_order[4];
    trlo_terrain_piece piece[4];
} trlo_ground_var3d;
void trlo_ground3d_init(trlo_ground_var3d* g, double spacing_x, double spacing_z, double world_scale,
                        const double bound_min[3], const double bound_max[3]);
void trlo_ground3d_update(trlo_ground_var3d* g, const double bound_min[3], const double bound_max[3]);
void trlo_ground3d_view(const trlo_ground_var3d* g, trlo_ground* out);

/* ---- batched drivers (loop over envs [env_begin, env_end); thread-safe on disjoint ranges) ---- */
typedef struct trlo_step_io {
    int num_envs;
    double dt;
    /* inputs, env-major */
    const double* pose;       /* [E][n] */
    const double* vel;        /* [E][n] */
    const double* tar_pose;   /* [E][n] */
    const double* tar_vel;    /* [E][n] */
    const unsigned char* joint_active; /* [E][N] or NULL */
    const double* kin_time;   /* [E] */
    const double* origin_pos; /* [E][3] or NULL (zero) */
    const double* origin_rot; /* [E][4] or NULL (identity) */
    const double* body_pos;   /* [E][N][3] */
    const double* body_vel;   /* [E][N][3] */
    const double* phase;      /* [E] or NULL */
    const unsigned char* flip;/* [E] or NULL */
    const trlo_ground* grounds; /* [num_grounds] */
    const int* env_to_ground; /* [E] or NULL (env % num_grounds) */
    int num_grounds;
    /* outputs */
    double* tau;              /* [E][n], overwritten (tau = 0 + control force) */
    double* kin_pose;         /* [E][n] */
    double* kin_vel;          /* [E][n] */
    double* kin_body_pos;     /* [E][N][3] FK of kin_pose */
    double* state;            /* [E][F] */
} trlo_step_io;

void trlo_step_range(const trlo_model* m, const trlo_obs_cfg* cfg, const trlo_step_io* io,
                     int env_begin, int env_end);

#ifdef __cplusplus
}
#endif
#endif
