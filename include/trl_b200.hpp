// trl_b200.hpp -- thin C++ shim over the C ABI (trl_b200.h) with reference-shaped classes, so C++ callers of the
// reference (controllers, scenarios) can switch the numerical core of this path without touching their call sites'
// shape. Header-only; link with -ltrl_b200. std::vector<double> stands in for Eigen::VectorXd (Eigen is not a
// dependency of this library): VectorXd::data()/size() map 1:1.
//
//   trl::cCharModel          <-> cCharacter::Init + cKinTree::Load + cPDController::LoadParams + cMotion::Load
//   trl::cImpPDController    <-> cImpPDController (sim/ImpPDController.h:9-55) incl. its cRBDModel
//                                (GetMassMat / GetBiasForce, sim/RBDModel.h:15-29), batched over E envs
//   trl::cRBDUtil            <-> cRBDUtil::BuildJacobian / BuildCOMJacobian / CalcGravityForce (sim/RBDUtil.h)
//   trl::cSimCharacter       <-> cSimCharacter::ApplyControlForces + cJoint::ApplyTau (sim/SimCharacter.cpp:589-603)
//   trl::cKinCharacter       <-> cKinCharacter::Pose / SetOriginPos / SetOriginRot (anim/KinCharacter.h:15-45)
//   trl::cGround             <-> cGround::SampleHeight (sim/Ground.h:93-97)
//   trl::cTerrainRLCharController <-> ParseGround + BuildPoliState (sim/TerrainRLCharController.h:31-87)
//   trl::cExpPDController    <-> cExpPDController::UpdateControlForce (sim/ExpPDController.h:15-35)
//   trl::cGroundVar          <-> cGroundVar2D / cGroundVar3D Init + Update on generated terrain (sim/GroundVar2D.h, GroundVar3D.h)
//   trl::cBatchedSimAdapter  <-> cSimAdapter (simAdapter/SimAdapter.h:29-81, terrainRLAdapter.swig) for E characters,
//                                over a caller-supplied rigid-body world (trl::cWorldHook; Bullet in the reference)
// Errors: the reference returns bool + printf / asserts; here every call returns bool and LastError() has the text.
#pragma once

#include <cmath>
#include <cstddef>
#include <string>
#include <vector>

#include "trl_b200.h"

namespace trl {

inline std::string LastError() { return trl_last_error(); }

class cCharModel {
public:
    cCharModel() = default;
    ~cCharModel() { Clear(); }
    cCharModel(const cCharModel&) = delete;
    cCharModel& operator=(const cCharModel&) = delete;

    // cCharacter::Init(char_file) + cKinCharacter::LoadMotion(motion_file)
    bool Init(const std::string& char_file, const std::string& motion_file = "", const double* gravity = nullptr) {
        Clear();
        mModel = trl_model_load(char_file.c_str(), motion_file.empty() ? nullptr : motion_file.c_str(), gravity);
        return mModel != nullptr;
    }
    // cRBDModel::Init(joint_mat, body_defs, gravity) + pd_params (row-major flat matrices)
    bool Init(int num_joints, const std::vector<double>& joint_mat, const std::vector<double>& body_defs,
              const std::vector<double>& pd_params, const double* gravity = nullptr) {
        Clear();
        mModel = trl_model_create(num_joints, joint_mat.data(), body_defs.data(), pd_params.data(), 0, nullptr, 0, gravity);
        return mModel != nullptr;
    }
    void Clear() {
        if (mModel) trl_model_free(mModel);
        mModel = nullptr;
    }
    int GetNumJoints() const { int N = 0; trl_model_dims(mModel, &N, nullptr, nullptr); return N; }
    int GetNumDof() const { int n = 0; trl_model_dims(mModel, nullptr, &n, nullptr); return n; }
    trl_model* Handle() const { return mModel; }

private:
    trl_model* mModel = nullptr;
};

// One batch of E characters on one GPU. The controller-shaped classes below share it (like the reference's
// controllers share one cSimCharacter / cRBDModel).
class cBatch {
public:
    cBatch() = default;
    ~cBatch() { Clear(); }
    cBatch(const cBatch&) = delete;
    cBatch& operator=(const cBatch&) = delete;
    bool Init(const cCharModel& model, int num_envs, int device, const trl_obs_cfg* cfg = nullptr) {
        Clear();
        mBatch = trl_batch_create(model.Handle(), num_envs, device, cfg);
        mNumEnvs = num_envs;
        mNumDof = model.GetNumDof();
        mNumJoints = model.GetNumJoints();
        return mBatch != nullptr;
    }
    void Clear() {
        if (mBatch) trl_batch_destroy(mBatch);
        mBatch = nullptr;
    }
    trl_batch* Handle() const { return mBatch; }
    int GetNumEnvs() const { return mNumEnvs; }
    int GetNumDof() const { return mNumDof; }
    int GetNumJoints() const { return mNumJoints; }

private:
    trl_batch* mBatch = nullptr;
    int mNumEnvs = 0, mNumDof = 0, mNumJoints = 0;
};

class cImpPDController {
public:
    // cImpPDController::Init(character, pd_params, gravity): gains come from the model's "PDControllers"
    void Init(cBatch* batch) {
        mBatch = batch;
        const size_t E = batch->GetNumEnvs(), n = batch->GetNumDof(), N = batch->GetNumJoints();
        mTarPose.assign(E * n, 0.0);
        mTarVel.assign(E * n, 0.0);
        mActive.assign(E * N, 1);
        mKp.clear();
        mKd.clear();
        mMass.clear();
        mBias.clear();
    }
    // cExpPDController::SetTargetTheta(joint_id, theta) / SetTargetVel, per env; `offset` = cKinTree::GetParamOffset
    void SetTargetTheta(int env, int param_offset, const std::vector<double>& theta) {
        for (size_t i = 0; i < theta.size(); ++i) mTarPose[(size_t)env * mBatch->GetNumDof() + param_offset + i] = theta[i];
    }
    void SetTargetVel(int env, int param_offset, const std::vector<double>& vel) {
        for (size_t i = 0; i < vel.size(); ++i) mTarVel[(size_t)env * mBatch->GetNumDof() + param_offset + i] = vel[i];
    }
    // GetPDCtrl(joint_id).SetActive(active)
    void SetActive(int env, int joint_id, bool active) { mActive[(size_t)env * mBatch->GetNumJoints() + joint_id] = active ? 1 : 0; }
    // SetKp / SetKd(joint_id, k): per-env gain overrides (allocated on first use from `defaults` [N])
    void SetKp(int env, int joint_id, double kp, const std::vector<double>& defaults) { Override(mKp, defaults)[(size_t)env * mBatch->GetNumJoints() + joint_id] = kp; }
    void SetKd(int env, int joint_id, double kd, const std::vector<double>& defaults) { Override(mKd, defaults)[(size_t)env * mBatch->GetNumJoints() + joint_id] = kd; }

    // UpdateControlForce(time_step, out_tau): out_tau [E*n] is accumulated (+=) like the reference.
    bool UpdateControlForce(double time_step, const std::vector<double>& pose, const std::vector<double>& vel,
                            std::vector<double>& out_tau, bool keep_rbd = false) {
        const size_t E = mBatch->GetNumEnvs(), n = mBatch->GetNumDof();
        if (out_tau.size() != E * n) out_tau.assign(E * n, 0.0);
        if (keep_rbd) { mMass.resize(E * n * n); mBias.resize(E * n); }
        const int rc = trl_imp_pd_update_control_force(
            mBatch->Handle(), time_step, pose.data(), vel.data(), mTarPose.data(), mTarVel.data(), mActive.data(),
            mKp.empty() ? nullptr : mKp.data(), mKd.empty() ? nullptr : mKd.data(), out_tau.data(),
            keep_rbd ? mMass.data() : nullptr, keep_rbd ? mBias.data() : nullptr);
        return rc == TRL_OK;
    }
    // cRBDModel::GetMassMat() / GetBiasForce() of the last UpdateControlForce(..., keep_rbd = true); [E][n][n], [E][n]
    const std::vector<double>& GetMassMat() const { return mMass; }
    const std::vector<double>& GetBiasForce() const { return mBias; }

private:
    std::vector<double>& Override(std::vector<double>& v, const std::vector<double>& defaults) {
        if (v.empty()) {
            const size_t E = mBatch->GetNumEnvs(), N = mBatch->GetNumJoints();
            v.resize(E * N);
            for (size_t e = 0; e < E; ++e)
                for (size_t j = 0; j < N; ++j) v[e * N + j] = defaults[j];
        }
        return v;
    }
    cBatch* mBatch = nullptr;
    std::vector<double> mTarPose, mTarVel, mKp, mKd, mMass, mBias;
    std::vector<unsigned char> mActive;
};

// cRBDUtil::BuildJacobian / BuildCOMJacobian / CalcGravityForce (sim/RBDUtil.h) for all E envs of a batch:
// J, Jcom [E][6][n] row-major, tau_g [E][n]. cRBDUtil::ExtractEndEffJacobian(joint_mat, J, joint_id) is a column mask
// of J (the blocks of the joints on the chain root..joint_id) and stays on the caller's side.
class cRBDUtil {
public:
    static bool BuildJacobian(cBatch& batch, const std::vector<double>& pose, std::vector<double>& out_J) {
        out_J.resize((size_t)batch.GetNumEnvs() * 6 * batch.GetNumDof());
        return trl_rbd_extras(batch.Handle(), pose.data(), out_J.data(), nullptr, nullptr) == TRL_OK;
    }
    static bool BuildCOMJacobian(cBatch& batch, const std::vector<double>& pose, std::vector<double>& out_J) {
        out_J.resize((size_t)batch.GetNumEnvs() * 6 * batch.GetNumDof());
        return trl_rbd_extras(batch.Handle(), pose.data(), nullptr, out_J.data(), nullptr) == TRL_OK;
    }
    static bool CalcGravityForce(cBatch& batch, const std::vector<double>& pose, std::vector<double>& out_g_force) {
        out_g_force.resize((size_t)batch.GetNumEnvs() * batch.GetNumDof());
        return trl_rbd_extras(batch.Handle(), pose.data(), nullptr, nullptr, out_g_force.data()) == TRL_OK;
    }
};

// cSimCharacter::ApplyControlForces(tau) + cJoint::ApplyTau for all E envs: clamped generalized torques [E][n] and the
// world-frame (torque, force) each body receives [E][N][6].
class cSimCharacter {
public:
    static bool ApplyControlForces(cBatch& batch, const std::vector<double>& pose, const std::vector<double>& tau,
                                   std::vector<double>& out_tau, std::vector<double>& out_body_wrench) {
        out_tau.resize((size_t)batch.GetNumEnvs() * batch.GetNumDof());
        out_body_wrench.resize((size_t)batch.GetNumEnvs() * batch.GetNumJoints() * 6);
        return trl_apply_control_forces(batch.Handle(), pose.data(), tau.data(), out_tau.data(), out_body_wrench.data()) == TRL_OK;
    }
};

class cKinCharacter {
public:
    void Init(cBatch* batch) {
        mBatch = batch;
        const size_t E = batch->GetNumEnvs();
        mTime.assign(E, 0.0);
        mOrigin.assign(E * 3, 0.0);
        mOriginRot.assign(E * 4, 0.0);
        for (size_t e = 0; e < E; ++e) mOriginRot[4 * e] = 1.0;
    }
    void SetTime(int env, double t) { mTime[env] = t; }
    void SetOriginPos(int env, double x, double y, double z) { mOrigin[3 * env] = x; mOrigin[3 * env + 1] = y; mOrigin[3 * env + 2] = z; }
    void SetOriginRot(int env, double w, double x, double y, double z) {
        mOriginRot[4 * env] = w; mOriginRot[4 * env + 1] = x; mOriginRot[4 * env + 2] = y; mOriginRot[4 * env + 3] = z;
    }
    // Update(dt): mTime += dt; Pose(mTime) -> GetPose() / GetVel()
    bool Update(double dt) {
        for (double& t : mTime) t += dt;
        const size_t E = mBatch->GetNumEnvs(), n = mBatch->GetNumDof();
        mPose.resize(E * n);
        mVel.resize(E * n);
        return trl_kin_char_pose(mBatch->Handle(), mTime.data(), mOrigin.data(), mOriginRot.data(), mPose.data(), mVel.data()) == TRL_OK;
    }
    const std::vector<double>& GetPose() const { return mPose; }
    const std::vector<double>& GetVel() const { return mVel; }

private:
    cBatch* mBatch = nullptr;
    std::vector<double> mTime, mOrigin, mOriginRot, mPose, mVel;
};

class cGround {
public:
    void Init(cBatch* batch) { mBatch = batch; }
    // SampleHeight(const Eigen::MatrixXd& pos /*rows x 3 per env*/, Eigen::VectorXd& out_h)
    bool SampleHeight(const std::vector<double>& pos, int samples_per_env, std::vector<double>& out_h,
                      std::vector<unsigned char>* out_valid = nullptr) {
        const size_t E = mBatch->GetNumEnvs();
        out_h.resize(E * samples_per_env);
        if (out_valid) out_valid->resize(E * samples_per_env);
        return trl_ground_sample_height(mBatch->Handle(), pos.data(), samples_per_env, out_h.data(),
                                        out_valid ? out_valid->data() : nullptr, nullptr) == TRL_OK;
    }

private:
    cBatch* mBatch = nullptr;
};

class cTerrainRLCharController {
public:
    void Init(cBatch* batch) { mBatch = batch; }
    int GetPoliStateSize() const { return trl_poli_state_size(mBatch->Handle()); }
    // ParseGround() + BuildPoliState(out_state); body_pos / body_vel [E][N][3] come from the physics engine
    bool BuildPoliState(const std::vector<double>& pose, const std::vector<double>& vel, const std::vector<double>& body_pos,
                        const std::vector<double>& body_vel, std::vector<double>& out_state,
                        const std::vector<double>* phase = nullptr, const std::vector<unsigned char>* flip = nullptr) {
        out_state.resize((size_t)mBatch->GetNumEnvs() * GetPoliStateSize());
        return trl_build_poli_state(mBatch->Handle(), pose.data(), vel.data(), body_pos.data(), body_vel.data(), nullptr,
                                    nullptr, nullptr, phase ? phase->data() : nullptr, flip ? flip->data() : nullptr,
                                    out_state.data(), nullptr) == TRL_OK;
    }

private:
    cBatch* mBatch = nullptr;
};

// cExpPDController-shaped explicit PD (sim/ExpPDController.h:15-35): same targets / gains as cImpPDController above.
class cExpPDController {
public:
    void Init(cBatch* batch) { mBatch = batch; }
    // UpdateControlForce(time_step, out_tau): tau [E][n] is accumulated like the reference's out_tau
    bool UpdateControlForce(double time_step, const std::vector<double>& pose, const std::vector<double>& vel,
                            const std::vector<double>& tar_pose, const std::vector<double>* tar_vel,
                            std::vector<double>& inout_tau, const std::vector<unsigned char>* active = nullptr) {
        inout_tau.resize((size_t)mBatch->GetNumEnvs() * mBatch->GetNumDof(), 0.0);
        return trl_exp_pd_update_control_force(mBatch->Handle(), time_step, pose.data(), vel.data(), tar_pose.data(),
                                               tar_vel ? tar_vel->data() : nullptr, active ? active->data() : nullptr, nullptr,
                                               nullptr, inout_tau.data()) == TRL_OK;
    }

private:
    cBatch* mBatch = nullptr;
};

// cGroundVar2D / cGroundVar3D-shaped procedural ground (sim/GroundVar2D.h, sim/GroundVar3D.h): Init builds one terrain per
// field from a terrain parameter file, Update scrolls them under the characters (cGround::Update).
class cGroundVar {
public:
    void Init(cBatch* batch) { mBatch = batch; }
    bool LoadParams(const std::string& terrain_file, double blend = 0.0, double world_scale = 4.0) {
        if (trl_terrain_load(terrain_file.c_str(), blend, &mCfg) != TRL_OK) return false;
        mCfg.world_scale = world_scale;
        return true;
    }
    trl_terrain_cfg& Params() { return mCfg; }
    bool Build(int num_fields, const std::vector<unsigned long long>& seeds, const std::vector<double>& bound_min,
               const std::vector<double>& bound_max) {
        mFields = num_fields;
        return trl_ground_generate(mBatch->Handle(), &mCfg, num_fields, seeds.data(), bound_min.data(), bound_max.data(), nullptr) == TRL_OK;
    }
    bool Update(const std::vector<double>& bound_min, const std::vector<double>& bound_max, std::vector<int>* out_changed = nullptr) {
        if (out_changed) out_changed->resize((size_t)mFields);
        return trl_ground_update(mBatch->Handle(), bound_min.data(), bound_max.data(), out_changed ? out_changed->data() : nullptr) == TRL_OK;
    }

private:
    cBatch* mBatch = nullptr;
    trl_terrain_cfg mCfg{};
    int mFields = 0;
};

// The rigid-body world behind the adapter (Bullet in the reference: cWorld + cSimCharacter's bodies). It stays the caller's:
// the adapter reads the generalized state and the body states from it and hands it the torques.
class cWorldHook {
public:
    virtual ~cWorldHook() = default;
    virtual void Reset(int num_envs) = 0;                                              // cScenarioSimChar::Reset
    // cSimCharacter::BuildPose / BuildVel [E][n], cSimObj::GetPos / GetLinearVelocity [E][N][3]
    virtual void GetState(std::vector<double>& pose, std::vector<double>& vel, std::vector<double>& body_pos,
                          std::vector<double>& body_vel) = 0;
    virtual void ApplyTau(const std::vector<double>& tau, double dt) = 0;              // ApplyControlForces + cWorld::Update(dt)
    virtual void HasFallen(std::vector<unsigned char>& out_fallen) = 0;                // cSimCharacter::HasFallen
    virtual double KinTime(int env) const { (void)env; return 0.0; }                   // clock of the kinematic reference
};

// cSimAdapter (simAdapter/SimAdapter.h:29-81, terrainRLAdapter.swig:1-44) for E characters at once: the same method names;
// every std::vector<double> of the reference is the env-major concatenation over the E envs.
class cBatchedSimAdapter {
public:
    // args: the reference's argv ("-arg_file=", "<file>", "-relative_file_path=", "<dir>/")
    cBatchedSimAdapter(const std::vector<std::string>& args, int num_envs, cWorldHook* world, int device = 0)
        : mArgs(args), mNumEnvs(num_envs), mDevice(device), mWorld(world) {}

    bool init() {  // SimAdapter.cpp:40-253
        std::string arg_file, rel;
        for (size_t i = 0; i + 1 < mArgs.size(); ++i) {
            if (mArgs[i].rfind("-arg_file=", 0) == 0) arg_file = mArgs[i + 1];
            if (mArgs[i].rfind("-relative_file_path=", 0) == 0) rel = mArgs[i + 1];
        }
        if (arg_file.empty()) return false;
        auto arg = [&](const char* key, const std::string& dflt) {
            char buf[1024];
            return trl_arg_file_get(arg_file.c_str(), key, buf, (int)sizeof(buf)) == TRL_OK ? std::string(buf) : dflt;
        };
        auto path = [&](const std::string& p) { return (p.empty() || p[0] == '/') ? p : rel + p; };
        const std::string char_file = path(arg("character_file", "")), motion_file = path(arg("motion_file", ""));
        if (char_file.empty() || !mModel.Init(char_file, motion_file)) return false;
        mNumUpdateSteps = std::stoi(arg("num_update_steps", "20"));
        const std::string ctrl = arg("char_ctrl", "");
        trl_obs_cfg cfg{TRL_OBS_TERRAIN_RL, TRL_SAMPLER_LINE, 200, 0, -0.5, 10.0, -1};
        if (ctrl.rfind("ct", 0) == 0) cfg = trl_obs_cfg{TRL_OBS_CT, TRL_SAMPLER_LINE, 0, 0, -0.5, 10.0, ctrl.find("phase") != std::string::npos ? 4 : -1};
        else if (ctrl.rfind("biped3D", 0) == 0) cfg = trl_obs_cfg{TRL_OBS_TERRAIN_RL, TRL_SAMPLER_GRID, 256, 16, -1.0, 2.0, -1};
        else if (ctrl.find("hopper") != std::string::npos) cfg.obs_kind = TRL_OBS_TERRAIN_RL_HOPPER;
        mHasPhase = cfg.num_phase_bins >= 0;
        if (!mBatch.Init(mModel, mNumEnvs, mDevice, &cfg)) return false;
        mGround.Init(&mBatch);
        const std::string terrain = path(arg("terrain_file", ""));
        if (terrain.empty()) trl_terrain_default_cfg(&mGround.Params());
        else if (!mGround.LoadParams(terrain, 0.0, std::stod(arg("world_scale", "1")))) return false;
        const int F = mNumEnvs < 4096 ? mNumEnvs : 4096;
        const bool is3d = mGround.Params().type == TRL_TERRAIN_VAR3D_FLAT;
        const double half = is3d ? 10.0 : 0.5 * mGround.Params().ground_width, shift = is3d ? 0.0 : -1.0;
        std::vector<unsigned long long> seeds((size_t)F);
        std::vector<double> bmin((size_t)F * 3, 0.0), bmax((size_t)F * 3, 0.0);
        for (int f = 0; f < F; ++f) {
            seeds[f] = (unsigned long long)mSeed + 104729ULL * (unsigned long long)f + 1ULL;
            bmin[3 * f] = -half + shift; bmax[3 * f] = half + shift;
            if (is3d) { bmin[3 * f + 2] = -half; bmax[3 * f + 2] = half; }
        }
        if (!mGround.Build(F, seeds, bmin, bmax)) return false;
        mPD.Init(&mBatch);
        mCtrl.Init(&mBatch);
        const size_t En = (size_t)mNumEnvs * mModel.GetNumDof();
        mTarPose.assign(En, 0.0);
        mTarVel.assign(En, 0.0);
        initEpoch();
        return true;
    }
    void initEpoch() {  // :255-260
        mWorld->Reset(mNumEnvs);
        mFrame = 0;
        Pull();
        mTarPose = mPose;  // default action: hold the current pose
        const int n = mModel.GetNumDof();
        for (int e = 0; e < mNumEnvs; ++e)
            for (int i = 0; i < 7; ++i) mTarPose[(size_t)e * n + i] = 0.0;
    }
    bool update() {  // :262-268 + cScenarioSimChar::Update: num_update_steps control substeps per animation frame
        const double dt = (1.0 / 30.0) / mNumUpdateSteps;
        for (int s = 0; s < mNumUpdateSteps; ++s) {
            Pull();
            mTau.assign(mPose.size(), 0.0);  // cSimCharacter::ClearJointTorques
            if (trl_imp_pd_update_control_force(mBatch.Handle(), dt, mPose.data(), mVel.data(), mTarPose.data(), mTarVel.data(),
                                                nullptr, nullptr, nullptr, mTau.data(), nullptr, nullptr) != TRL_OK)
                return false;
            mWorld->ApplyTau(mTau, dt);
        }
        ++mFrame;
        return true;
    }
    bool updateAction(const std::vector<double>& action) {  // :441-456: PD targets of the actuated dofs, [E][n - 7]
        const int n = mModel.GetNumDof(), a = n - 7;
        if (action.size() != (size_t)mNumEnvs * a) return false;
        for (int e = 0; e < mNumEnvs; ++e)
            for (int i = 0; i < a; ++i) mTarPose[(size_t)e * n + 7 + i] = action[(size_t)e * a + i];
        return true;
    }
    void handleUpdatedAction() {}
    std::vector<double> getState() {  // :493-508: ParseGround + BuildPoliState, [E][F]
        Pull();
        std::vector<double> state, phase;
        const std::vector<double>* ph = nullptr;
        if (mHasPhase) { phase.assign((size_t)mNumEnvs, std::fmod(mFrame * (1.0 / 30.0), 1.0)); ph = &phase; }
        mCtrl.BuildPoliState(mPose, mVel, mBodyPos, mBodyVel, state, ph);
        return state;
    }
    std::vector<double> calcReward() {  // cScenarioExpImitate::CalcReward vs the kinematic reference, [E]
        Pull();
        const int n = mModel.GetNumDof();
        std::vector<double> t((size_t)mNumEnvs), kin_pose((size_t)mNumEnvs * n), kin_vel((size_t)mNumEnvs * n), reward((size_t)mNumEnvs, 0.0);
        for (int e = 0; e < mNumEnvs; ++e) t[e] = mWorld->KinTime(e);
        std::vector<unsigned char> fallen;
        mWorld->HasFallen(fallen);
        if (trl_kin_char_pose(mBatch.Handle(), t.data(), nullptr, nullptr, kin_pose.data(), kin_vel.data()) != TRL_OK) return reward;
        trl_imitate_calc_reward(mBatch.Handle(), mPose.data(), mVel.data(), kin_pose.data(), kin_vel.data(), mBodyVel.data(),
                                fallen.data(), reward.data(), nullptr);
        return reward;
    }
    std::vector<unsigned char> agentHasFallen() {  // :458-465, [E]
        std::vector<unsigned char> f;
        mWorld->HasFallen(f);
        return f;
    }
    bool needUpdatedAction() const { return true; }                       // :467-475 (policy query rate = 30 Hz)
    size_t getActionSpaceSize() const { return (size_t)mModel.GetNumDof() - 7; }   // :477-483
    size_t getObservationSpaceSize() const { return (size_t)trl_poli_state_size(mBatch.Handle()); }  // :485-491
    void setRandomSeed(int seed) { mSeed = seed; }                         // :510-515
    void finish() { mBatch.Clear(); mModel.Clear(); }
    const std::vector<double>& tau() const { return mTau; }
    cBatch& batch() { return mBatch; }

private:
    void Pull() { mWorld->GetState(mPose, mVel, mBodyPos, mBodyVel); }
    std::vector<std::string> mArgs;
    int mNumEnvs, mDevice, mNumUpdateSteps = 20, mFrame = 0, mSeed = 0;
    bool mHasPhase = false;
    cWorldHook* mWorld;
    cCharModel mModel;
    cBatch mBatch;
    cGroundVar mGround;
    cImpPDController mPD;
    cTerrainRLCharController mCtrl;
    std::vector<double> mPose, mVel, mBodyPos, mBodyVel, mTarPose, mTarVel, mTau;
};

}  // namespace trl
