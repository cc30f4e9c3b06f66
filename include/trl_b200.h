/*
 * trl_b200.h -- C ABI of the B200-native batched control-and-observation step
 * (drop-in for TerrainRLSim's cImpPDController / cRBDModel / cKinTree / cKinCharacter /
 *  cGround / cTerrainRLCharController numerical core, batched over E independent characters).
 *
 * The reference has no plugin ABI: the path sits behind C++ virtual classes and the SWIG class
 * cSimAdapter. Each entry point below names the reference interface it
 * replaces (file:line in the reference tree). Semantics, argument meaning and error behaviour
 * follow the reference; batching adds a leading [E] dimension to every vector.
 *
 * Conventions
 *  - all arrays are double, env-major contiguous ([E][n], [E][N][3] ...), unless noted;
 *  - pose/vel layout is the reference's: root (x y z | qw qx qy qz) then joints in index order,
 *    vel root (vx vy vz | wx wy wz 0), spherical (wx wy wz 0)          (sim/SimCharacter.cpp:1037-1115);
 *  - every function returns 0 on success or a negative trl_status; it never aborts
 *    (the reference's assert()/printf+bool convention mapped to codes; message via trl_last_error());
 *  - pointer mode (like cuBLAS): TRL_PTR_HOST  -> array arguments are host pointers, the call stages
 *    H2D, runs the kernels, stages D2H and returns when the results are in the host arrays;
 *    TRL_PTR_DEVICE -> array arguments are device pointers on the batch's device, the call only
 *    enqueues kernels on the batch's stream (asynchronous; trl_batch_sync() to wait);
 *  - one trl_batch per calling thread / stream (thread-compatible, like one scenario per thread in
 *    scenarios/ScenarioTrain.cpp:176-191); a trl_model is immutable and may be shared;
 *  - there is NO CPU fallback: creating a batch without a usable CUDA device fails with
 *    TRL_ERR_NO_DEVICE.
 */
#ifndef TRL_B200_H
#define TRL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TRL_VERSION 100

#define TRL_JOINT_COLS 19 /* cKinTree::eJointDescMax, anim/KinTree.h:24-46 */
#define TRL_BODY_COLS 17  /* cKinTree::eBodyParamMax, anim/KinTree.h:57-77 */
#define TRL_PD_COLS 18    /* cPDController::eParamMax, sim/PDController.h:12-32 */
#define TRL_MAX_JOINTS 32
#define TRL_MAX_DOF 64

typedef enum trl_status {
    TRL_OK = 0,
    TRL_ERR_INVALID_ARG = -1,
    TRL_ERR_IO = -2,          /* file missing / unreadable */
    TRL_ERR_PARSE = -3,       /* JSON / arg-file syntax or schema */
    TRL_ERR_UNSUPPORTED = -4, /* joint / shape type outside the scoped path, model too large */
    TRL_ERR_NO_DEVICE = -5,   /* no CUDA device: there is no CPU fallback */
    TRL_ERR_CUDA = -6,
    TRL_ERR_NO_MOTION = -7,
    TRL_ERR_NO_GROUND = -8
} trl_status;

typedef struct trl_model trl_model;
typedef struct trl_batch trl_batch;

/* ------------------------------------------------------------------------------------------
 * model  (cCharacter::Init + cKinTree::Load/LoadBodyDefs + cPDController::LoadParams + cMotion::Load)
 * ---------------------------------------------------------------------------------------- */

/* Parse the reference's character JSON ("Skeleton"."Joints", "BodyDefs", "PDControllers";
 * anim/KinTree.cpp:127-199,451-499,980-1041, sim/PDController.cpp:29-91) and optional motion JSON
 * ("Loop", "Frames"; anim/Motion.cpp:36-65,147-205). gravity NULL = (0,-9.8,0) (util/MathUtil.h:21).
 * Returns NULL on failure (see trl_last_error()). */
trl_model* trl_model_load(const char* character_json_path, const char* motion_json_path_or_null,
                          const double* gravity3_or_null);

/* Same from flat matrices: mirrors cRBDModel::Init(joint_mat, body_defs, gravity) (sim/RBDModel.cpp:16-37)
 * and cImpPDController::Init(..., pd_params, gravity) (sim/ImpPDController.cpp:26-39).
 * frames [F][1+n] with col 0 = frame DURATION as stored on disk (anim/Motion.cpp:207-217), or NULL. */
trl_model* trl_model_create(int num_joints, const double* joint_mat, const double* body_defs,
                            const double* pd_params, int num_frames, const double* frames_or_null, int loop,
                            const double* gravity3_or_null);

/* cKinTree::GetNumJoints / GetNumDof (anim/KinTree.cpp:622-632), cMotion::GetNumFrames */
int trl_model_dims(const trl_model* m, int* num_joints, int* num_dof, int* num_frames);
/* copies of the parsed matrices ([N][19], [N][17], [N][18]); any output may be NULL */
int trl_model_get_matrices(const trl_model* m, double* joint_mat, double* body_defs, double* pd_params);
/* frames [F][1+n] with col 0 = cumulative start time (post cMotion::PostProcessFrames); returns loop flag in *loop */
int trl_model_get_motion(const trl_model* m, double* frames, int* loop);
void trl_model_free(trl_model* m);

/* Arg files: "-key= value" tokens, // comments (util/ArgParser.cpp:1-344). Looks up one key; returns 0 and
 * copies the value string, or TRL_ERR_INVALID_ARG when the key is absent. */
int trl_arg_file_get(const char* arg_file_path, const char* key, char* out_value, int out_capacity);

/* Character state files (data/states/ *.txt, {"Pose": [n], "Vel": [n]}): cCharacter::ReadState / WriteState
 * (anim/Character.cpp:277-342,370-379). Reading normalises the root's and the spherical joints' quaternions
 * (cKinTree::PoseProcessPose); a key the file does not hold leaves the caller's array untouched. Writing uses the
 * reference's number format (std::to_string). pose / vel [n]. */
int trl_state_read(const trl_model* m, const char* state_file, double* inout_pose, double* inout_vel);
int trl_state_write(const trl_model* m, const char* state_file, const double* pose, const double* vel);

/* ------------------------------------------------------------------------------------------
 * batch
 * ---------------------------------------------------------------------------------------- */
typedef enum trl_obs_kind {
    TRL_OBS_TERRAIN_RL = 0,        /* cTerrainRLCharController features (sim/TerrainRLCharController.cpp:344-436) */
    TRL_OBS_TERRAIN_RL_HOPPER = 1, /* + cMonopedHopperController overrides (sim/MonopedHopperController.cpp:1533-1552) */
    TRL_OBS_CT = 2,                /* cCtController 2-D features (sim/CtController.cpp:507-632) */
    TRL_OBS_CT_3D = 3              /* cCtController 3-D features, ENABLE_MAX_STATE (sim/CtController.h:6) */
} trl_obs_kind;

typedef enum trl_sampler { TRL_SAMPLER_LINE = 0, TRL_SAMPLER_GRID = 1 } trl_sampler;

typedef struct trl_obs_cfg {
    int obs_kind;       /* trl_obs_kind */
    int sampler;        /* LINE: CalcGroundSamplePos of cTerrainRLCharController (:310-319);
                           GRID: cCtController / cBipedController3D res x res square (CtController.cpp:189-228,
                           BipedController3D.cpp:1927-1965) */
    int num_samples;    /* G = GetNumGroundSamples() (grid: res*res) */
    int grid_res;
    double view_min;    /* LINE: gViewMin (-0.5); GRID: min view dist */
    double view_max;    /* LINE: view dist (10);  GRID: max view dist */
    int num_phase_bins; /* <0: no phase block; >=0: cCtPhaseController block of 1+bins entries (CtPhaseController.cpp:66-139) */
} trl_obs_cfg;

typedef enum trl_pointer_mode { TRL_PTR_HOST = 0, TRL_PTR_DEVICE = 1 } trl_pointer_mode;

/* device >= 0: CUDA device ordinal. Fails with TRL_ERR_NO_DEVICE (returns NULL) if CUDA is unusable. */
trl_batch* trl_batch_create(const trl_model* m, int num_envs, int device, const trl_obs_cfg* cfg);
void trl_batch_destroy(trl_batch* b);
int trl_batch_set_pointer_mode(trl_batch* b, int mode);
/* Use an existing cudaStream_t (e.g. torch's current stream). As in CUDA, NULL is the legacy default stream;
 * TRL_STREAM_OWN switches back to the batch's own non-blocking stream (the initial state). Kernels enqueued by
 * this library can be captured into a CUDA graph on that stream (DEVICE pointer mode only). */
#define TRL_STREAM_OWN ((void*)(~(unsigned long long)0))
int trl_batch_set_stream(trl_batch* b, void* cuda_stream);
int trl_batch_sync(trl_batch* b);
/* HOST pointer mode only: with `on`, trl_step_fused returns as soon as the step's uploads, kernels and downloads are
 * enqueued; the outputs are valid after trl_batch_sync(). The input and output arrays must be page-locked and must not
 * be touched, and no other call may be made on this batch, until then. Two batches driven alternately this way overlap
 * one's download with the other's upload and kernels (the synchronous call is PCIe-bound). */
int trl_batch_set_host_async(trl_batch* b, int on);
/* GetPoliStateSize() (sim/TerrainRLCharController.cpp:438-493, CtController.cpp:482-504) */
int trl_poli_state_size(const trl_batch* b);
int trl_batch_num_envs(const trl_batch* b);
/* number of this library's kernel launches issued through this batch so far */
long long trl_batch_launch_count(const trl_batch* b);

/* Diagnostics: kernel timeline of everything launched while enabled (process-wide, current device). `out32` receives
 * 16 pairs (first block start, last block end) in GPU-timer nanoseconds -- slot 0 k_imp_pd, 1 k_pd_tree, 2 k_pd_h,
 * 3 k_pd_solve, 4 k_kin_thread, 5 k_obs_env, 6 k_obs_samples; untouched slots hold (2^64-1, 0). Reading re-arms it. */
int trl_debug_trace_enable(int on);
int trl_debug_trace_read(unsigned long long* out32);

/* ------------------------------------------------------------------------------------------
 * (1) stable PD
 *   cRBDModel::Update(pose, vel)                         sim/RBDModel.cpp:39-55
 *   cImpPDController::UpdateControlForce(dt, out_tau)    sim/ImpPDController.cpp:48-74,137-196
 *   targets as set by SetTargetTheta/SetTargetVel        sim/ExpPDController.cpp:70-92 (root slots ignored)
 *   GetPDCtrl(j).SetActive(bool)                          -> joint_active [E][N] (NULL = all active)
 *   SetKp/SetKd(joint, k)                                 -> kp/kd [E][N] overrides (NULL = "PDControllers" gains)
 *   inout_tau [E][n] is ACCUMULATED (+=) like the reference's out_tau
 *   GetMassMat()/GetBiasForce()                           -> optional out_mass [E][n][n], out_bias [E][n]
 * ---------------------------------------------------------------------------------------- */
int trl_imp_pd_update_control_force(trl_batch* b, double dt, const double* pose, const double* vel,
                                    const double* tar_pose, const double* tar_vel,
                                    const unsigned char* joint_active_or_null,
                                    const double* kp_or_null, const double* kd_or_null, double* inout_tau,
                                    double* out_mass_or_null, double* out_bias_or_null);

/* Explicit PD: cExpPDController::UpdateControlForce -> cPDController::UpdateControlForce / CalcJointTau{Revolute,
 * Prismatic, Spherical} (sim/ExpPDController.cpp:59-66,148-159; sim/PDController.cpp:134-149,302-428), with the same target
 * post-processing as the stable-PD entry (normalised spherical targets, world-coordinate targets of UseWorldCoord joints,
 * sim/PDController.cpp:191-273). Per active valid joint: revolute / prismatic tau += kp (tar - theta) + kd (tar_vel - vel);
 * spherical tau += kp * axis-angle(q^-1 * tar) + kd (tar_vel - vel) on slots 0..2; planar and fixed joints add nothing.
 * Same arguments as trl_imp_pd_update_control_force (no M / C outputs); inout_tau [E][n] is ACCUMULATED; dt <= 0 is a
 * no-op like the reference (ExpPDController.cpp:62). */
int trl_exp_pd_update_control_force(trl_batch* b, double dt, const double* pose, const double* vel,
                                    const double* tar_pose, const double* tar_vel,
                                    const unsigned char* joint_active_or_null,
                                    const double* kp_or_null, const double* kd_or_null, double* inout_tau);

/* cSimCharacter::ApplyControlForces + cJoint::ApplyTau (sim/SimCharacter.cpp:589-603; sim/Joint.cpp:162-221,300-318,
 * 536-562): the step right after the torque computation -- per joint, the generalized torques become a (torque, force)
 * pair, ClampTotalTorque / ClampTotalForce limit their norms (joint_mat TorqueLim / ForceLim), and the joint's world
 * frame (FK of `pose`) maps them to equal-and-opposite world torques / forces on the child and parent bodies.
 * pose, tau [E][n]; out_tau [E][n] = clamped generalized torques (in place allowed: out_tau may equal tau in DEVICE
 * mode); out_body_wrench [E][N][6] = (torque xyz, force xyz) per body in world coordinates. Outputs may be NULL. */
int trl_apply_control_forces(trl_batch* b, const double* pose, const double* tau, double* out_tau, double* out_body_wrench);

/* cRBDUtil::BuildJacobian / BuildCOMJacobian / CalcGravityForce (sim/RBDUtil.cpp:256-275,294-345,937-982): the RBD
 * quantities the FSM controllers read every substep next to M and C (BipedController3D.cpp:1065-1072,1301-1388,
 * DogController.cpp:897,982). pose [E][n]; out_J, out_Jcom [E][6][n] row-major (rows = wx wy wz vx vy vz in world
 * coordinates; columns of dead dof slots are zero); out_gravity_force [E][n]. Any output may be NULL. */
int trl_rbd_extras(trl_batch* b, const double* pose, double* out_J, double* out_Jcom, double* out_gravity_force);

/* cRBDUtil::BuildEndEffectorJacobian (sim/RBDUtil.cpp:181-233), read every substep by the FSM controllers for the stance /
 * swing foot (sim/BipedController3D.cpp:1301-1388, sim/DogController.cpp:972-1060): the world-coordinate Jacobian of joint
 * `joint_id`'s frame -- the columns of the joints on the chain root .. joint_id, zero elsewhere.
 * joint_ids [E] (one joint per env: the stance leg differs between envs), or NULL to use `joint_id_all` for every env.
 * pose [E][n]; out_J [E][6][n] row-major like trl_rbd_extras. A joint id outside [0, N) is TRL_ERR_INVALID_ARG (checked for
 * joint_id_all; out-of-range entries of joint_ids give an all-zero Jacobian for that env). */
int trl_rbd_end_effector_jacobian(trl_batch* b, const double* pose, const int* joint_ids_or_null, int joint_id_all,
                                  double* out_J);

/* ------------------------------------------------------------------------------------------
 * (2) kinematics
 *   cKinTree::JointWorldTrans for every joint (anim/KinTree.cpp:1099-1112) -> out_joint_world [E][N][12]
 *     (row-major 3x4 [R | p]);  cKinTree::CalcBodyPartPos (:299-308) -> out_body_pos [E][N][3]
 *   cKinCharacter::Pose(t) = CalcPose(t) + finite-difference CalcVel (anim/KinCharacter.cpp:111-123,280-325),
 *     cMotion::CalcFrame / CalcIndexPhase (anim/Motion.cpp:130-139,240-273), cKinTree::LerpPoses (:1529-1569);
 *     origin_pos [E][3] / origin_rot [E][4] (w,x,y,z) = SetOriginPos / SetOriginRot, NULL = zero / identity
 * ---------------------------------------------------------------------------------------- */
int trl_kin_tree_fk(trl_batch* b, const double* pose, double* out_joint_world_or_null, double* out_body_pos_or_null);
int trl_kin_char_pose(trl_batch* b, const double* time, const double* origin_pos_or_null,
                      const double* origin_rot_or_null, double* out_pose, double* out_vel_or_null);

/* ------------------------------------------------------------------------------------------
 * (3) ground
 *   Heightfields are registered once (always HOST pointers) and kept resident in HBM.
 *   kind: 1 = cGroundVar2D segments (2 per field), 2 = cGroundVar3D slabs (4 per field), 3 = plane (h = 0).
 *   data: concatenated float heights exactly as the reference stores them (already multiplied by world scale,
 *   sim/GroundVar2D.cpp:451-466, GroundVar3D.cpp:449-455); data_offset [F][P] element offsets; res [F][P][2] =
 *   (grid width x, grid length z); origin [F][P][3] = cSimObj::GetPos(); scale [F][P][3] = GetScaling();
 *   bounds [F][P][4] = (min_x, max_x, min_z, max_z) used by ContainsPos (2-D: the Bullet AABB in x,
 *   GroundVar2D.cpp:697-700; 3-D: mMin/mMax, GroundVar3D.cpp:596-600).
 *   env_to_field [E] (NULL: env % num_fields).
 *   cGround::SampleHeight(pos, valid) scalar overload numerics (GroundVar2D.cpp:109-124,621-638;
 *   GroundVar3D.cpp:154-170,631-684): pos [E][S][3] -> out_h [E][S], out_valid [E][S], out_cell [E][S][3] =
 *   (piece, i, j) for the bit-exact index check (piece = -1: no piece contains pos, h = -inf).
 * ---------------------------------------------------------------------------------------- */
int trl_ground_set_heightfields(trl_batch* b, int kind, int num_fields, int pieces_per_field, const float* data,
                                long long data_count, const long long* data_offset, const int* res,
                                const double* origin, const double* scale, const double* bounds,
                                const int* env_to_field_or_null);
int trl_ground_sample_height(trl_batch* b, const double* pos, int samples_per_env, double* out_h,
                             unsigned char* out_valid_or_null, int* out_cell_or_null);

/* ------------------------------------------------------------------------------------------
 * (3b) procedural terrain on the device: the terrains are generated and scrolled where the
 *   sampling kernels read them, one per field, with the reference's random stream (std::default_random_engine order).
 *   cGroundVar2D::Init / Update / BuildSegment   sim/GroundVar2D.cpp:33-96,261-450   (two segments, cTerrainGen2D::Build*
 *                                                sim/TerrainGen2D.cpp:127-653)
 *   cGroundVar3D::Init / Update / BuildSlab      sim/GroundVar3D.cpp:24-141,350-420  (four slabs, cTerrainGen3D::BuildFlat)
 *   cGroundFactory::ParseParamsJson              sim/GroundFactory.cpp:11-60 + cTerrainGen2D::LoadParams (data/terrain/ *.txt)
 * ---------------------------------------------------------------------------------------- */
#define TRL_TERRAIN_NUM_PARAMS 40 /* cTerrainGen2D::eParamsMax, in the order of sim/TerrainGen2D.h:13-66 */
typedef enum trl_terrain_type { /* cGround::eType, sim/Ground.h:26-55 */
    TRL_TERRAIN_PLANE = 0, TRL_TERRAIN_VAR2D_FLAT = 1, TRL_TERRAIN_VAR2D_GAPS = 2, TRL_TERRAIN_VAR2D_STEPS = 3,
    TRL_TERRAIN_VAR2D_WALLS = 4, TRL_TERRAIN_VAR2D_BUMPS = 5, TRL_TERRAIN_VAR2D_MIXED = 6, TRL_TERRAIN_VAR2D_NARROW_GAPS = 7,
    TRL_TERRAIN_VAR2D_SLOPES = 8, TRL_TERRAIN_VAR2D_SLOPES_GAPS = 9, TRL_TERRAIN_VAR2D_SLOPES_WALLS = 10,
    TRL_TERRAIN_VAR2D_SLOPES_STEPS = 11, TRL_TERRAIN_VAR2D_SLOPES_MIXED = 12, TRL_TERRAIN_VAR2D_SLOPES_NARROW_GAPS = 13,
    TRL_TERRAIN_VAR2D_CLIFFS = 14, TRL_TERRAIN_VAR3D_FLAT = 15
} trl_terrain_type;

typedef struct trl_terrain_cfg {
    int type;                               /* trl_terrain_type */
    int pad;
    double params[TRL_TERRAIN_NUM_PARAMS];  /* blended parameter set (cGround::CalcBlendParams) */
    double ground_width;                    /* GroundWidth (20); var3d uses 40 like cGroundVar3D::Init */
    double world_scale;                     /* -world_scale= (4): heights are stored multiplied by it */
    double spacing_x, spacing_z;            /* VertSpacingX / Z (0.2): var3d vertex spacing */
} trl_terrain_cfg;

/* Defaults of cGround::tParams and cTerrainGen2D::gParamDefs (type var2d_flat, width 20, spacing 0.2, world scale 4). */
int trl_terrain_default_cfg(trl_terrain_cfg* out);
/* Parses a data/terrain/ *.txt file ("Type", "GroundWidth", "VertSpacingX/Z", "Params": [ {..}, .. ]); `blend` selects /
 * interpolates between the parameter sets like cGround::SetParamBlend. Types other than plane / var2d_* / var3d_flat are
 * TRL_ERR_UNSUPPORTED. world_scale is left at its default (it comes from the arg file). */
int trl_terrain_load(const char* terrain_file, double blend, trl_terrain_cfg* out);
/* Generates `num_fields` terrains on the device and registers them like trl_ground_set_heightfields (which this call
 * replaces). seeds [F] = cGround::tParams::mRandSeed per terrain; bound_min / bound_max [F][3] = the view bounds passed to
 * cGround::Init (var2d reads x only). Plane terrains register the flat ground. HOST pointers regardless of the mode. */
int trl_ground_generate(trl_batch* b, const trl_terrain_cfg* cfg, int num_fields, const unsigned long long* seeds,
                        const double* bound_min, const double* bound_max, const int* env_to_field_or_null);
/* cGround::Update(dt, bound_min, bound_max) for every terrain: a segment / slab the view bounds have reached is rebuilt
 * (continuing the terrain's random stream), the others are kept. bound_min / bound_max [F][3] and out_changed_or_null [F]
 * (1 = rebuilt, 2 = a rebuilt piece did not fit its slot: TRL_ERR_UNSUPPORTED in HOST mode) follow the pointer mode; in DEVICE mode the call only enqueues the kernel. */
int trl_ground_update(trl_batch* b, const double* bound_min, const double* bound_max, int* out_changed_or_null);
/* Reads terrain `field` back in the layout of trl_ground_set_heightfields (pieces in the reference's scan order; var2d
 * rows are replicated to the reference's 3): out_data holds the pieces back to back (capacity in floats), res [P][2],
 * origin / scale [P][3], bounds [P][4], data_offset [P]. Returns the number of floats written or a negative status. */
long long trl_ground_read_field(trl_batch* b, int field, float* out_data, long long capacity, long long* data_offset,
                                int* res, double* origin, double* scale, double* bounds);

/* ------------------------------------------------------------------------------------------
 * (4) observation
 *   cTerrainRLCharController::ParseGround + BuildPoliState (sim/TerrainRLCharController.cpp:281-436) and the
 *   cCtController / cCtPhaseController / hopper / biped3D variants selected by trl_obs_cfg.
 *   body_pos / body_vel [E][N][3] = cSimObj::GetPos / GetLinearVelocity of each body part;
 *   body_quat [E][N][4] (w,x,y,z) / body_angvel [E][N][3] only for TRL_OBS_CT_3D;
 *   ground_vel [E][3] (NULL = 0); phase [E] (NULL = 0); flip [E] stance flip (NULL = 0; Q12);
 *   vel [E][n] only read for TRL_OBS_TERRAIN_RL_HOPPER (may be NULL otherwise).
 *   out_state [E][F], F = trl_poli_state_size(); out_cell_or_null [E][G][3] ground-sample cells.
 * ---------------------------------------------------------------------------------------- */
int trl_build_poli_state(trl_batch* b, const double* pose, const double* vel_or_null, const double* body_pos,
                         const double* body_vel, const double* body_quat_or_null, const double* body_angvel_or_null,
                         const double* ground_vel_or_null, const double* phase_or_null,
                         const unsigned char* flip_or_null, double* out_state, int* out_cell_or_null);

/* ------------------------------------------------------------------------------------------
 * fused step: the measured entry point. One env-step = (1) stable-PD torque, (2) kin-char pose+vel,
 * FK of the kin pose (body positions), (3)+(4) ground features + packed policy state.
 * tau is OVERWRITTEN (cSimCharacter::Update clears joint torques first, sim/SimCharacter.cpp:95-108).
 * Any output pointer may be NULL to skip that part.
 * ---------------------------------------------------------------------------------------- */
typedef struct trl_step_args {
    double dt;
    const double* pose;                /* [E][n] */
    const double* vel;                 /* [E][n] */
    const double* tar_pose;            /* [E][n] */
    const double* tar_vel;             /* [E][n] */
    const unsigned char* joint_active; /* [E][N] or NULL */
    const double* kin_time;            /* [E] */
    const double* origin_pos;          /* [E][3] or NULL */
    const double* origin_rot;          /* [E][4] or NULL */
    const double* body_pos;            /* [E][N][3] */
    const double* body_vel;            /* [E][N][3] */
    const double* phase;               /* [E] or NULL */
    const unsigned char* flip;         /* [E] or NULL */
    double* out_tau;                   /* [E][n] */
    double* out_kin_pose;              /* [E][n] */
    double* out_kin_vel;               /* [E][n] */
    double* out_kin_body_pos;          /* [E][N][3] */
    double* out_state;                 /* [E][F] */
} trl_step_args;

int trl_step_fused(trl_batch* b, const trl_step_args* args);

/* Extended fused step: fewer bytes across PCIe for host-coupled callers. All members optional (NULL).
 *   out_reward [E]      imitation reward epilogue (trl_imitate_calc_reward on the step's own kin pose / vel, which then
 *                       need not be requested: out_kin_* may be NULL and stay on the device). Needs a motion, kin_time
 *                       and body_vel. cScenarioExpImitate::CalcReward, scenarios/ScenarioExpImitate.cpp:13-159.
 *   fallen [E]          per-env "has fallen" flag for the reward (NULL = none).
 *   out_state_f32 [E][F] single-precision copy of the policy state (the reference feeds the state to Caffe as floats);
 *                       args->out_state may then be NULL.
 * In trl_step_args, tar_vel == NULL means zero target velocities (the reference's default), and out_state == NULL skips
 * the observation: a caller that queries the policy every k-th substep (cCtController::UpdateCalcTau,
 * sim/CtController.cpp:300-325: 1 in 20) passes out_state only on those. */
typedef struct trl_step_opts {
    double* out_reward;
    const unsigned char* fallen;
    float* out_state_f32;
} trl_step_opts;

int trl_step_fused_ex(trl_batch* b, const trl_step_args* args, const trl_step_opts* opts_or_null);

/* ------------------------------------------------------------------------------------------
 * imitation reward: cScenarioExpImitate::CalcReward (scenarios/ScenarioExpImitate.cpp:13-159)
 *   with cKinTree::NormalizePoseHeading / CalcPoseErr / CalcVelErr (anim/KinTree.cpp:1319-1426,1647-1672),
 *   cRBDUtil::CalcCoM / CalcCoMVel (sim/RBDUtil.cpp:499-598), cSimCharacter::CalcCOMVel (sim/SimCharacter.cpp:359-377).
 *   pose / vel [E][n] = simulated character, kin_pose / kin_vel [E][n] = kinematic reference (trl_kin_char_pose),
 *   body_vel [E][N][3] = body linear velocities, fallen [E] (NULL = none fallen; a fallen character scores 0).
 *   End-effector positions of the simulated character are the FK of `pose` (the reference reads them from Bullet).
 *   out_reward [E]; out_terms [E][5] = pose, vel, end-effector, root, com errors (optional).
 * ---------------------------------------------------------------------------------------- */
int trl_imitate_calc_reward(trl_batch* b, const double* pose, const double* vel, const double* kin_pose,
                            const double* kin_vel, const double* body_vel, const unsigned char* fallen_or_null,
                            double* out_reward, double* out_terms_or_null);

/* per-env status word of the last stable-PD call: bit0 non-finite tau, bit1 non-positive pivot
 * (the reference would assert). out [E] ints, follows the pointer mode. */
int trl_batch_get_status(trl_batch* b, int* out_status);

/* Measurement utility: achieved FP64 TFLOP/s of a register-resident DFMA loop on `device` (the FP64 roofline
 * denominator reported by bench.py; MEASURED_PEAKS.json carries no FP64 figure). */
int trl_device_fp64_probe(int device, double* out_tflops);

const char* trl_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
