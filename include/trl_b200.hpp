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
// Errors: the reference returns bool + printf / asserts; here every call returns bool and LastError() has the text.
#pragma once

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

}  // namespace trl
