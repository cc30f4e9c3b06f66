// Internal definitions shared by the host runtime and the CUDA kernels.
#pragma once

#include <string>
#include <vector>

#include "../../include/trl_b200.h"

// joint types, anim/KinTree.h:13-22
enum { JT_REVOLUTE = 0, JT_PLANAR = 1, JT_PRISMATIC = 2, JT_FIXED = 3, JT_SPHERICAL = 4, JT_NONE = 5 };

// live dof column kinds (how a column of the joint subspace S looks in the root frame),
// cRBDUtil::BuildJointSubspace*, sim/RBDUtil.cpp:733-828
enum { CK_ANG = 0,      // omega = R_j[:,axis], v = p_j x omega   (revolute z, spherical x/y/z, planar theta, root ang)
       CK_LIN = 1,      // omega = 0, v = R_j[:,axis]             (prismatic x, planar x/y)
       CK_ROOT_LIN = 2  // omega = 0, v = row `axis` of R(root quat) (root translation, RBDUtil.cpp:776-783)
};

// Integer index tables of a model as signed bytes (every value is < 64 or -1): 848 B, staged in shared memory by
// every kernel so the tree walks never wait on global / L2 loads (with ~200 KB of shared memory per SM carved out
// for the per-env workspaces, L1 is too small to keep them resident).
struct alignas(16) ModelIdx {
    signed char parent[TRL_MAX_JOINTS];
    signed char type[TRL_MAX_JOINTS];
    signed char offset[TRL_MAX_JOINTS];
    signed char size[TRL_MAX_JOINTS];
    signed char valid_body[TRL_MAX_JOINTS];
    signed char pd_valid[TRL_MAX_JOINTS];
    signed char level[TRL_MAX_JOINTS];
    signed char first_col[TRL_MAX_JOINTS];
    signed char ncol[TRL_MAX_JOINTS];
    signed char level_joint[TRL_MAX_JOINTS];
    signed char slot_base[TRL_MAX_JOINTS];  // pose-feature slot of body j, cTerrainRLCharController packing (bodies 1..N-1)
    signed char slot_ct[TRL_MAX_JOINTS];    // same for cCtController packing (bodies 0..N-1); -1 = invalid body
    signed char end_eff[TRL_MAX_JOINTS];    // cKinTree::IsEndEffector
    signed char level_start[TRL_MAX_JOINTS + 16];
    signed char dfs_event[2 * TRL_MAX_JOINTS];  // depth-first walk: j = enter joint j, -1 - j = leave joint j (2N events)
    signed char col_joint[TRL_MAX_DOF];
    signed char col_kind[TRL_MAX_DOF];
    signed char col_axis[TRL_MAX_DOF];
    signed char col_dof[TRL_MAX_DOF];     // index into pose/vel vectors
    signed char col_parent[TRL_MAX_DOF];  // Featherstone's lambda() expanded to dof columns, -1 at the top
    signed char dof_col[TRL_MAX_DOF];     // dof -> live column or -1
    signed char dof_joint[TRL_MAX_DOF];
};
static_assert(sizeof(ModelIdx) % 16 == 0, "ModelIdx is copied with 16-byte loads");

// Index tables of the block-cooperative stable-PD kernel (k_pd_block): the branch-sparse storage of H = M + dt*Kd.
// Column c only couples with the columns on its own ancestor chain (Featherstone's branch-induced sparsity), so row c
// is stored as [H(c,c), H(c,lambda c), H(c,lambda^2 c), ...] starting at rowbase[c]; the entry for an ancestor a of c
// sits at rowbase[c] + cdepth[c] - cdepth[a].
struct alignas(16) PdBlkIdx {
    short rowbase[TRL_MAX_DOF];
    signed char cdepth[TRL_MAX_DOF];         // number of column-ancestors of c
    signed char cd_order[TRL_MAX_DOF];       // columns sorted by cdepth (back substitution runs depth by depth)
    signed char cd_start[TRL_MAX_DOF + 16];  // start of each depth in cd_order, [max_cdepth + 2] entries
    signed char use_world[TRL_MAX_JOINTS];   // pd_params UseWorldCoord of the joint (sim/PDController.h:31)
    short aba_off[TRL_MAX_JOINTS];           // k_pd_aba: first scratch slot of joint j (1 + 14 k doubles for k live dofs)
    signed char preorder[TRL_MAX_JOINTS];    // joints in depth-first pre-order (the order ModelIdx::dfs_event enters them)
    // k_pd_aba works branch by branch (a branch = the subtree of one child of the root), one warp per branch:
    signed char nchild[TRL_MAX_JOINTS];      // number of children of joint j
    signed char b_ev0[TRL_MAX_JOINTS];       // branch b: first event in ModelIdx::dfs_event (its events are contiguous)
    signed char b_ev1[TRL_MAX_JOINTS];       //           one past its last event
    signed char b_pre0[TRL_MAX_JOINTS];      //           first joint in `preorder`
    signed char b_pre1[TRL_MAX_JOINTS];      //           one past its last joint
    signed char b_depth[TRL_MAX_JOINTS];     //           deepest level in it (its level slots come first in its area)
    short xoff[TRL_MAX_JOINTS];              // joints with >= 2 children: offset (elements, within the branch's area) of
                                             // their [V6 A6 IA21 P6] block; -1 otherwise
    int num_branches;
    int branch_area;                         // elements per branch area: 22 per level below the root + the blocks above
    int pad_[2];
    int hcount;                              // entries of the sparse H
    int max_cdepth;
    int aba_slots;                           // scratch doubles per env of k_pd_aba
    int all_valid;                           // every body valid (k_pd_aba's precondition)
};
static_assert(sizeof(PdBlkIdx) % 16 == 0, "PdBlkIdx is copied with 16-byte loads");

// Flattened, immutable character model as the kernels read it (resident in HBM, ~5 KB).
struct DevModel {
    int N;           // joints
    int n;           // dofs incl. dead slots (cKinTree::GetNumDof)
    int nl;          // live dof columns (dead: root slot 6, spherical slot 3)
    int num_levels;  // tree depth levels
    int num_frames;  // motion frames (0 = none)
    int loop;
    int has_spherical;
    int pad0;
    double gravity[3];
    double motion_duration;
    double cycle_delta[3];  // cKinCharacter::CalcCycleRootDelta, anim/KinCharacter.cpp:264-277

    ModelIdx ix;  // small integer tables; every kernel stages them in shared memory
    PdBlkIdx bx;  // sparse-H index tables of k_pd_block
    double attR[TRL_MAX_JOINTS][9];  // cKinTree::BuildAttachTrans rotation (Rz*Ry*Rx of AttachTheta), row-major
    double attT[TRL_MAX_JOINTS][3];  // attach point
    double mass[TRL_MAX_JOINTS];
    double com[TRL_MAX_JOINTS][3];    // body attach point = CoM in joint coordinates (RBDUtil.cpp:675-684)
    double idiag[TRL_MAX_JOINTS][3];  // BuildMomentInertia{Box,Capsule} diagonal (RBDUtil.cpp:623-673)
    double kp[TRL_MAX_JOINTS];
    double kd[TRL_MAX_JOINTS];
    double joint_w[TRL_MAX_JOINTS];  // DiffWeight / sum|DiffWeight| (cScenarioExpImitate::CalcJointWeights)
    // per-joint constants of the stable-PD kernels, packed [N][21]: attach rotation (9), attach point (3), mass, CoM (3),
    // inertia diagonal (3), Kp, Kd -- copied to shared memory as one block
    alignas(16) double pd_cst[TRL_MAX_JOINTS * 21];
    double torque_lim[TRL_MAX_JOINTS];  // joint_mat TorqueLim / ForceLim (cJoint::ClampTotalTorque / ClampTotalForce)
    double force_lim[TRL_MAX_JOINTS];
    double total_mass;

};

struct trl_model {
    int N = 0, n = 0;
    std::vector<double> joint_mat, body_defs, pd_params;
    std::vector<double> frames;  // [F][1+n], col 0 cumulative time
    int num_frames = 0;
    int loop = 0;
    double gravity[3] = {0, -9.8, 0};
    DevModel dev;
};

// per-piece terrain descriptor in HBM (160 bytes). The device copy of a heightfield is re-pitched at registration: every
// piece starts on a 16-byte boundary and its rows are `pitch` floats apart (res_x rounded up to a multiple of 4), so a
// row segment that starts at a column multiple of 4 can be moved with one bulk asynchronous copy (k_obs_samples).
struct DevPiece {
    long long data_offset;  // first float of the piece in the device array (multiple of 4)
    int res_x, res_z;
    int pitch;              // floats between consecutive rows on the device
    int pad_[3];
    double origin[3];
    double scale[3];
    double bounds[4];
    double rscale[2];  // correctly-rounded 1/scale_x, 1/scale_z (host IEEE division) for div_by_const
    double half[2];    // (res_x - 1) * 0.5, (res_z - 1) * 0.5 (exact products)
    double cmax3[2];   // 3-D clamp bounds res_x - 1.0 - tol, res_z - 1.0 - tol (GroundVar3D.cpp:602-621), host-evaluated
};
static_assert(sizeof(DevPiece) == 160, "DevPiece is copied with 16-byte loads");

struct DevGround {
    int kind;  // 0 none, 1 var2d, 2 var3d, 3 plane
    int num_fields;
    int pieces_per_field;
    // k_obs_samples stages the cell rectangle an env's sample window covers in shared memory: up to patch_nbuf rectangles
    // (one per terrain piece the window touches) of patch_h rows x patch_w floats each; patch_nbuf = 0: never stage
    int patch_w, patch_h, patch_nbuf;
    int table_ok;  // every entry of the batch's local sample table is 0 or in [2^-300, 2^300] (see k_obs_env)
    int pad_;
    const float* data;
    const DevPiece* pieces;   // [F][P]
    const int* env_to_field;  // [E] or nullptr
};

// procedural terrain (trl_terrain.cu)
struct TerrainGenCfg {
    int type;
    int pieces;             // 2 (var2d segments) or 4 (var3d slabs)
    double params[TRL_TERRAIN_NUM_PARAMS];
    double ground_width, world_scale, spacing_x, spacing_z;
    long long slot_floats;  // capacity of one piece's slot in the device heightfield array (multiple of 4)
};
struct TerrainState {       // per terrain, device-resident
    unsigned long long rand_x;  // position in the terrain's minstd_rand0 stream
    int flip;                   // cGroundVar2D::mFlipSeg
    int order[4];               // cGroundVar3D::mSlabOrder
    int count[4];               // vertices of piece id (0 = empty)
    int overflow;               // a piece did not fit its slot
    int pad_[2];
    double aabb[4][4];          // rigid-body AABB of piece id: min x, max x, min z, max z
    double mn[4][3], mx[4][3];  // tSlab::mMin / mMax
    DevPiece piece[4];          // descriptors by piece id (the registered array holds them in scan order)
};
int trl_launch_terrain(const TerrainGenCfg& cfg, int F, const unsigned long long* d_seeds, const double* d_bound_min,
                       const double* d_bound_max, TerrainState* d_states, float* d_data, DevPiece* d_pieces, int* d_changed,
                       int init, void* stream);

void trl_set_error(const std::string& msg);
int trl_build_dev_model(trl_model* m);  // fills m->dev; returns status

// kernel launchers (trl_kernels.cu); all enqueue on `stream` and return a cudaError_t as int
struct StepPtrs {
    double dt;
    const double *pose, *vel, *tar_pose, *tar_vel;
    const unsigned char* joint_active;
    const double *kp, *kd;
    double* tau;
    int tau_accumulate;
    double *out_mass, *out_bias;
    int* status;
};

int trl_trace_enable(int on);
int trl_trace_read(unsigned long long* out32);
size_t trl_pd_scratch_doubles(const DevModel& hm);
int trl_pd_scratch_tile();
// returns the number of kernels launched (> 0) or a negative cudaError_t
int trl_launch_imp_pd(const DevModel* dm, const DevModel& hm, int E, const StepPtrs& p, double* scratch, void* stream);
int trl_launch_rbd_extras(const DevModel* dm, const DevModel& hm, int E, const double* pose, double* out_J,
                          double* out_Jcom, double* out_gforce, void* stream);
int trl_launch_apply_tau(const DevModel* dm, const DevModel& hm, int E, const double* pose, const double* tau,
                         double* out_tau, double* out_body_wrench, void* stream);
int trl_launch_fk(const DevModel* dm, const DevModel& hm, int E, const double* pose, double* out_joint_world,
                  double* out_body_pos, void* stream);
int trl_launch_kin_char(const DevModel* dm, const DevModel& hm, const double* frames, int E, const double* time,
                        const double* origin_pos, const double* origin_rot, double* out_pose, double* out_vel,
                        void* stream);
int trl_launch_kin(const DevModel* dm, const DevModel& hm, const double* frames, int E, const double* time,
                   const double* origin_pos, const double* origin_rot, double* out_pose, double* out_vel,
                   double* out_body_pos, void* stream);
int trl_launch_imitate_reward(const DevModel* dm, const DevModel& hm, const DevGround& g, int E, const double* pose0,
                              const double* vel0, const double* pose1, const double* vel1, const double* body_vel,
                              const unsigned char* fallen, double* out_reward, double* out_terms, void* stream);
int trl_launch_sample_height(const DevGround& g, int E, const double* pos, int S, double* out_h,
                             unsigned char* out_valid, int* out_cell, void* stream);
int trl_launch_poli_state(const DevModel* dm, const DevModel& hm, const DevGround& g, const trl_obs_cfg& cfg, int F,
                          int E, const double* pose, const double* vel, const double* body_pos,
                          const double* body_vel, const double* body_quat, const double* body_angvel,
                          const double* ground_vel, const double* phase, const unsigned char* flip,
                          const double* sample_local, void* env_rec, double* out_state, int* out_cell, void* stream,
                          void* side_stream = nullptr, void* ev_fork = nullptr, void* ev_join = nullptr);
int trl_launch_exp_pd(const DevModel* dm, const DevModel& hm, int E, const StepPtrs& p, void* stream);
int trl_launch_ee_mask(const DevModel* dm, int E, int n, const int* joint_ids, int joint_id_all, double* J, void* stream);
size_t trl_obs_env_rec_bytes();
int trl_launch_f64_to_f32(const double* src, float* dst, size_t count, void* stream);
