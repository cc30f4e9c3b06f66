// ref_wrap.cpp -- C entry points into the UNMODIFIED reference sources (TEST INFRASTRUCTURE).
//
// Linked with the reference's own util/MathUtil, util/Rand, util/FileUtil, util/JsonUtil, util/json/*,
// sim/SpAlg, sim/RBDUtil, sim/RBDModel, anim/KinTree, anim/Motion, anim/Character, anim/KinCharacter,
// anim/KinController, sim/Controller, sim/PDController, sim/ExpPDController, sim/ImpPDController, sim/Ground,
// sim/TerrainGen2D, sim/TerrainGen3D, sim/GroundVar2D, sim/GroundVar3D translation units, compiled where they lie
// under /root/reference against oracle/eigen_shim (Eigen stand-in) and oracle/ref_stubs (Bullet-side stand-ins)
// into oracle/_ref/libtrl_ref.so (recipe: oracle/Makefile, target `ref`).  This file only marshals flat arrays
// into the reference's classes and back; every number it returns is computed by reference code.
// Only tests/ and tools/make_golden.py load the library.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>

#include "anim/KinCharacter.h"
#include "anim/KinTree.h"
#include "anim/Motion.h"
#include "sim/GroundVar2D.h"
#include "sim/GroundVar3D.h"
#include "sim/ImpPDController.h"
#include "sim/RBDModel.h"
#include "sim/RBDUtil.h"
#include "sim/SimCharacter.h"
#include "util/FileUtil.h"
#include "util/MathUtil.h"

namespace {

Eigen::MatrixXd to_mat(const double* src, int rows, int cols)  // row-major flat -> MatrixXd
{
	Eigen::MatrixXd m(rows, cols);
	for (int i = 0; i < rows; ++i)
		for (int j = 0; j < cols; ++j) m(i, j) = src[i * cols + j];
	return m;
}
Eigen::VectorXd to_vec(const double* src, int n)
{
	Eigen::VectorXd v(n);
	for (int i = 0; i < n; ++i) v[i] = src[i];
	return v;
}
template <class M> void from_mat(const M& m, double* dst)  // -> row-major flat
{
	if (!dst) return;
	for (int i = 0; i < m.rows(); ++i)
		for (int j = 0; j < m.cols(); ++j) dst[i * m.cols() + j] = m(i, j);
}
template <class V> void from_vec(const V& v, double* dst, int n)
{
	if (!dst) return;
	for (int i = 0; i < n; ++i) dst[i] = v[i];
}

// protected statics of cRBDUtil
struct RBDProbe : public cRBDUtil
{
	using cRBDUtil::BuildInertiaSpatialMat;
};

struct Model
{
	Eigen::MatrixXd joint_mat, body_defs, pd_params;
	tVector gravity;
	int N, n;
	cRBDModel rbd;
	bool has_pd;
};

struct KinChar
{
	cKinCharacter kc;
};

// terrain classes expose their pieces through protected members
struct Var2DProbe : public cGroundVar2D
{
	const tSegment& Seg(int s) const { return *GetSegment(s); }
};
struct Var3DProbe : public cGroundVar3D
{
	const tSlab& Slab(int s) const { return static_cast<const cGroundVar3D*>(this)->GetSlab(s); }
};
struct Ground
{
	int kind;  // 1 = var2d, 2 = var3d
	std::shared_ptr<cWorld> world;
	std::shared_ptr<Var2DProbe> g2;
	std::shared_ptr<Var3DProbe> g3;
	cGround::tParams params;
	cGround* base() const { return kind == 1 ? static_cast<cGround*>(g2.get()) : static_cast<cGround*>(g3.get()); }
};

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------------------------
void* trlr_model_create(int N, const double* joint_mat, const double* body_defs, const double* pd_params, const double* gravity)
{
	Model* m = new Model();
	m->N = N;
	m->joint_mat = to_mat(joint_mat, N, cKinTree::eJointDescMax);
	m->body_defs = to_mat(body_defs, N, cKinTree::eBodyParamMax);
	m->has_pd = pd_params != nullptr;
	if (pd_params) m->pd_params = to_mat(pd_params, N, cPDController::eParamMax);
	m->gravity = tVector(gravity[0], gravity[1], gravity[2], 0);
	m->n = cKinTree::GetNumDof(m->joint_mat);
	m->rbd.Init(m->joint_mat, m->body_defs, m->gravity);
	return m;
}
void trlr_model_free(void* h) { delete static_cast<Model*>(h); }
int trlr_num_dof(void* h) { return static_cast<Model*>(h)->n; }

// cCharacter::LoadSkeleton / cKinTree::Load, cKinTree::LoadBodyDefs, cPDController::LoadParams on the reference's files.
// out arrays sized [N][19], [N][17], [N][18] (pass NULL to query N only); returns N or -1
int trlr_load_character(const char* char_file, double* out_joint_mat, double* out_body_defs, double* out_pd_params)
{
	std::ifstream f_stream(char_file);
	Json::Reader reader;
	Json::Value root;
	bool succ = reader.parse(f_stream, root);
	f_stream.close();
	if (!succ || root[cCharacter::gSkeletonKey].isNull()) return -1;
	Eigen::MatrixXd joint_mat;
	if (!cKinTree::Load(root[cCharacter::gSkeletonKey], joint_mat)) return -1;
	int N = static_cast<int>(joint_mat.rows());
	from_mat(joint_mat, out_joint_mat);
	if (out_body_defs)
	{
		Eigen::MatrixXd body_defs;
		if (!cKinTree::LoadBodyDefs(char_file, body_defs)) return -1;
		from_mat(body_defs, out_body_defs);
	}
	if (out_pd_params)
	{
		Eigen::MatrixXd pd_params;
		if (!cPDController::LoadParams(char_file, pd_params)) return -1;
		from_mat(pd_params, out_pd_params);
	}
	return N;
}

// ---------------------------------------------------------------------------------------------------------------
// rigid body dynamics: cRBDModel::Update -> GetMassMat / GetBiasForce, cRBDUtil::*
// ---------------------------------------------------------------------------------------------------------------
void trlr_rbd_update(void* h, const double* pose, const double* vel, double* out_mass, double* out_bias)
{
	Model* m = static_cast<Model*>(h);
	m->rbd.Update(to_vec(pose, m->n), to_vec(vel, m->n));
	from_mat(m->rbd.GetMassMat(), out_mass);
	from_vec(m->rbd.GetBiasForce(), out_bias, m->n);
}
void trlr_solve_inv_dyna(void* h, const double* pose, const double* vel, const double* acc, double* out_tau)
{
	Model* m = static_cast<Model*>(h);
	m->rbd.Update(to_vec(pose, m->n), to_vec(vel, m->n));
	Eigen::VectorXd tau;
	cRBDUtil::SolveInvDyna(m->rbd, to_vec(acc, m->n), tau);
	from_vec(tau, out_tau, m->n);
}
void trlr_inertia_spatial(void* h, int j, double* out36)
{
	Model* m = static_cast<Model*>(h);
	cSpAlg::tSpMat I = RBDProbe::BuildInertiaSpatialMat(m->body_defs, j);
	from_mat(I, out36);
}
void trlr_rbd_extras(void* h, const double* pose, double* out_J, double* out_Jcom, double* out_gravity_force)
{
	Model* m = static_cast<Model*>(h);
	m->rbd.Update(to_vec(pose, m->n), Eigen::VectorXd::Zero(m->n));
	Eigen::MatrixXd J, Jcom;
	cRBDUtil::BuildJacobian(m->rbd, J);
	from_mat(J, out_J);
	if (out_Jcom)
	{
		cRBDUtil::BuildCOMJacobian(m->rbd, J, Jcom);
		from_mat(Jcom, out_Jcom);
	}
	if (out_gravity_force)
	{
		Eigen::VectorXd g;
		cRBDUtil::CalcGravityForce(m->rbd, g);
		from_vec(g, out_gravity_force, m->n);
	}
}
void trlr_end_eff_jacobian(void* h, const double* pose, int joint_id, double* out_J)
{
	Model* m = static_cast<Model*>(h);
	m->rbd.Update(to_vec(pose, m->n), Eigen::VectorXd::Zero(m->n));
	Eigen::MatrixXd J;
	cRBDUtil::BuildEndEffectorJacobian(m->rbd, joint_id, J);
	from_mat(J, out_J);
}
void trlr_calc_com(void* h, const double* pose, const double* vel, double* com, double* com_vel)
{
	Model* m = static_cast<Model*>(h);
	tVector c, v;
	cRBDUtil::CalcCoM(m->joint_mat, m->body_defs, to_vec(pose, m->n), to_vec(vel, m->n), c, v);
	from_vec(c, com, 3);
	from_vec(v, com_vel, 3);
}

// ---------------------------------------------------------------------------------------------------------------
// kinematics: cKinTree::*
// ---------------------------------------------------------------------------------------------------------------
void trlr_child_parent_trans(void* h, const double* pose, int j, double* out16)
{
	Model* m = static_cast<Model*>(h);
	from_mat(cKinTree::ChildParentTrans(m->joint_mat, to_vec(pose, m->n), j), out16);
}
void trlr_joint_world_trans(void* h, const double* pose, int j, double* out16)
{
	Model* m = static_cast<Model*>(h);
	from_mat(cKinTree::JointWorldTrans(m->joint_mat, to_vec(pose, m->n), j), out16);
}
void trlr_body_part_pos(void* h, const double* pose, int j, double* out3)
{
	Model* m = static_cast<Model*>(h);
	from_vec(cKinTree::CalcBodyPartPos(m->joint_mat, m->body_defs, to_vec(pose, m->n), j), out3, 3);
}
void trlr_vel_to_pose_diff(void* h, const double* pose, const double* vel, double* out)
{
	Model* m = static_cast<Model*>(h);
	Eigen::VectorXd d;
	cKinTree::VelToPoseDiff(m->joint_mat, to_vec(pose, m->n), to_vec(vel, m->n), d);
	from_vec(d, out, m->n);
}
void trlr_post_process_pose(void* h, double* pose)
{
	Model* m = static_cast<Model*>(h);
	Eigen::VectorXd p = to_vec(pose, m->n);
	cKinTree::PoseProcessPose(m->joint_mat, p);
	from_vec(p, pose, m->n);
}
void trlr_calc_vel(void* h, const double* pose0, const double* pose1, double dt, double* out_vel)
{
	Model* m = static_cast<Model*>(h);
	Eigen::VectorXd v;
	cKinTree::CalcVel(m->joint_mat, to_vec(pose0, m->n), to_vec(pose1, m->n), dt, v);
	from_vec(v, out_vel, m->n);
}
void trlr_lerp_poses(void* h, const double* pose0, const double* pose1, double lerp, double* out)
{
	Model* m = static_cast<Model*>(h);
	Eigen::VectorXd p;
	cKinTree::LerpPoses(m->joint_mat, to_vec(pose0, m->n), to_vec(pose1, m->n), lerp, p);
	from_vec(p, out, m->n);
}
double trlr_calc_heading(void* h, const double* pose)
{
	Model* m = static_cast<Model*>(h);
	return cKinTree::CalcHeading(m->joint_mat, to_vec(pose, m->n));
}
void trlr_build_origin_trans(void* h, const double* pose, double* out16)
{
	Model* m = static_cast<Model*>(h);
	from_mat(cKinTree::BuildOriginTrans(m->joint_mat, to_vec(pose, m->n)), out16);
}
// cMathUtil::QuaternionToAxisAngle (w,x,y,z) -> axis[3], theta
void trlr_quat_to_axis_angle(const double* q, double* out_axis, double* out_theta)
{
	tVector axis;
	double theta;
	cMathUtil::QuaternionToAxisAngle(tQuaternion(q[0], q[1], q[2], q[3]), axis, theta);
	from_vec(axis, out_axis, 3);
	*out_theta = theta;
}

// ---------------------------------------------------------------------------------------------------------------
// PD controllers: the reference's cImpPDController / cExpPDController on a stand-in character
// (oracle/ref_stubs/sim/SimCharacter.h).  Targets enter through SetTargetTheta / SetTargetVel, gains through SetKp /
// SetKd, the active mask through GetPDCtrl(j).SetActive -- the calls a reference controller makes.
// ---------------------------------------------------------------------------------------------------------------
static void setup_pd(Model* m, cSimCharacter& character, cExpPDController& ctrl, const double* pose, const double* vel,
					 const double* tar_pose, const double* tar_vel, const unsigned char* joint_active, const double* kp_joint,
					 const double* kd_joint)
{
	for (int j = 0; j < m->N; ++j)
	{
		if (!ctrl.IsValidPDCtrl(j)) continue;
		int off = cKinTree::GetParamOffset(m->joint_mat, j);
		int size = cKinTree::GetParamSize(m->joint_mat, j);
		if (kp_joint) ctrl.SetKp(j, kp_joint[j]);
		if (kd_joint) ctrl.SetKd(j, kd_joint[j]);
		ctrl.SetTargetTheta(j, to_vec(tar_pose + off, size));
		ctrl.SetTargetVel(j, to_vec(tar_vel + off, size));
		if (joint_active) ctrl.GetPDCtrl(j).SetActive(joint_active[j] != 0);
	}
	(void)character;
	(void)pose;
	(void)vel;
}
int trlr_imp_pd_control_force(void* h, double dt, const double* pose, const double* vel, const double* tar_pose,
							  const double* tar_vel, const unsigned char* joint_active, const double* kp_joint,
							  const double* kd_joint, double* inout_tau, double* out_mass, double* out_bias)
{
	Model* m = static_cast<Model*>(h);
	if (!m->has_pd) return -1;
	cSimCharacter character;
	character.InitStub(m->joint_mat, m->body_defs);
	character.SetPose(to_vec(pose, m->n));
	character.SetVel(to_vec(vel, m->n));
	cImpPDController ctrl;
	ctrl.Init(&character, m->pd_params, m->gravity);  // owns its cRBDModel, updates it in UpdateControlForce
	setup_pd(m, character, ctrl, pose, vel, tar_pose, tar_vel, joint_active, kp_joint, kd_joint);
	Eigen::VectorXd tau = to_vec(inout_tau, m->n);
	ctrl.UpdateControlForce(dt, tau);
	from_vec(tau, inout_tau, m->n);
	if (out_mass || out_bias) trlr_rbd_update(h, pose, vel, out_mass, out_bias);
	return 0;
}
int trlr_exp_pd_control_force(void* h, double dt, const double* pose, const double* vel, const double* tar_pose,
							  const double* tar_vel, const unsigned char* joint_active, const double* kp_joint,
							  const double* kd_joint, double* inout_tau)
{
	Model* m = static_cast<Model*>(h);
	if (!m->has_pd) return -1;
	cSimCharacter character;
	character.InitStub(m->joint_mat, m->body_defs);
	character.SetPose(to_vec(pose, m->n));
	character.SetVel(to_vec(vel, m->n));
	cExpPDController ctrl;
	ctrl.Init(&character, m->pd_params);
	setup_pd(m, character, ctrl, pose, vel, tar_pose, tar_vel, joint_active, kp_joint, kd_joint);
	Eigen::VectorXd tau = to_vec(inout_tau, m->n);
	ctrl.UpdateControlForce(dt, tau);
	from_vec(tau, inout_tau, m->n);
	return 0;
}
int trlr_pd_valid(void* h, int j)
{
	Model* m = static_cast<Model*>(h);
	cSimCharacter character;
	character.InitStub(m->joint_mat, m->body_defs);
	return character.GetJoint(j).IsValid() ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// motion / kinematic character: cKinCharacter::Init(char_file, motion_file), Pose(t), cMotion::*
// ---------------------------------------------------------------------------------------------------------------
void* trlr_kin_char_create(const char* char_file, const char* motion_file)
{
	KinChar* k = new KinChar();
	if (!k->kc.Init(char_file, motion_file ? motion_file : ""))
	{
		delete k;
		return nullptr;
	}
	return k;
}
void trlr_kin_char_free(void* h) { delete static_cast<KinChar*>(h); }
void trlr_kin_char_dims(void* h, int* N, int* n, int* num_frames, int* loop, double* duration)
{
	KinChar* k = static_cast<KinChar*>(h);
	*N = k->kc.GetNumJoints();
	*n = k->kc.GetNumDof();
	*num_frames = k->kc.GetMotion().GetNumFrames();
	*loop = k->kc.GetMotion().IsLoop() ? 1 : 0;
	*duration = k->kc.GetMotionDuration();
}
// frames after cMotion::PostProcessFrames: [F][1 + n], column 0 = cumulative start time
void trlr_motion_frames(void* h, double* out)
{
	KinChar* k = static_cast<KinChar*>(h);
	const cMotion& mo = k->kc.GetMotion();
	int F = mo.GetNumFrames(), n = mo.GetNumDof();
	for (int f = 0; f < F; ++f)
	{
		out[f * (1 + n)] = mo.GetFrameTime(f);
		cMotion::tFrame fr = mo.GetFrame(f);
		for (int i = 0; i < n; ++i) out[f * (1 + n) + 1 + i] = fr[i];
	}
}
void trlr_motion_index_phase(void* h, double time, int* out_idx, double* out_phase)
{
	KinChar* k = static_cast<KinChar*>(h);
	k->kc.GetMotion().CalcIndexPhase(time, *out_idx, *out_phase);
}
// cKinCharacter::Pose(t): GetPose / GetVel afterwards
void trlr_kin_char_pose(void* h, double time, const double* origin_pos, const double* origin_rot, double* out_pose,
						double* out_vel)
{
	KinChar* k = static_cast<KinChar*>(h);
	tVector op = origin_pos ? tVector(origin_pos[0], origin_pos[1], origin_pos[2], 0) : tVector::Zero();
	tQuaternion orot = origin_rot ? tQuaternion(origin_rot[0], origin_rot[1], origin_rot[2], origin_rot[3]) : tQuaternion::Identity();
	k->kc.SetOriginPos(op);
	k->kc.SetOriginRot(orot);
	k->kc.Pose(time);
	int n = k->kc.GetNumDof();
	from_vec(k->kc.GetPose(), out_pose, n);
	from_vec(k->kc.GetVel(), out_vel, n);
}
// cCharacter::ReadState(state_file) -> pose, vel
int trlr_read_state(void* h, const char* state_file, double* out_pose, double* out_vel)
{
	KinChar* k = static_cast<KinChar*>(h);
	if (!k->kc.ReadState(state_file)) return -1;
	int n = k->kc.GetNumDof();
	from_vec(k->kc.GetPose(), out_pose, n);
	from_vec(k->kc.GetVel(), out_vel, n);
	return n;
}

// ---------------------------------------------------------------------------------------------------------------
// terrain: cGroundVar2D / cGroundVar3D built the way cGroundFactory does (sim/GroundFactory.cpp:11-61,149-181)
// ---------------------------------------------------------------------------------------------------------------
static bool parse_ground_params(const char* param_file, cGround::tParams& out_params)
{
	std::ifstream f_stream(param_file);
	Json::Reader reader;
	Json::Value root;
	bool succ = reader.parse(f_stream, root);
	f_stream.close();
	if (!succ) return false;
	if (!root[cGround::gTypeKey].isNull())
	{
		cGround::ParseType(root[cGround::gTypeKey].asString(), out_params.mType);
	}
	out_params.mGroundWidth = root.get(cGround::gGroundWidthKey, out_params.mGroundWidth).asDouble();
	out_params.mVertSpacingX = root.get(cGround::gVertSpacingXKey, out_params.mVertSpacingX).asDouble();
	out_params.mVertSpacingZ = root.get(cGround::gVertSpacingZKey, out_params.mVertSpacingZ).asDouble();
	if (!root[cGround::gParamsKey].isNull())
	{
		Json::Value params_arr = root[cGround::gParamsKey];
		int num = params_arr.size();
		out_params.mParamArr.resize(0, 0);
		cGround::eClass ground_class = cGround::GetClassFromType(out_params.mType);
		for (int i = 0; i < num; ++i)
		{
			Eigen::VectorXd curr_params;
			if (ground_class == cGround::eClassVar2D) succ &= cGroundVar2D::ParseParamsJson(params_arr.get(i, 0), curr_params);
			else if (ground_class == cGround::eClassVar3D) succ &= cGroundVar3D::ParseParamsJson(params_arr.get(i, 0), curr_params);
			else return false;
			if (succ)
			{
				if (i == 0) out_params.mParamArr.resize(num, curr_params.size());
				out_params.mParamArr.row(i) = curr_params;
			}
		}
	}
	return succ;
}
void* trlr_ground_create(const char* param_file, double world_scale, unsigned long seed, double blend)
{
	Ground* g = new Ground();
	if (!parse_ground_params(param_file, g->params))
	{
		delete g;
		return nullptr;
	}
	g->params.mHasRandSeed = true;
	g->params.mRandSeed = seed;
	g->params.mBlend = blend;
	g->world = std::shared_ptr<cWorld>(new cWorld(world_scale));
	cGround::eClass ground_class = cGround::GetClassFromType(g->params.mType);
	double half_width = g->params.mGroundWidth / 2;
	if (ground_class == cGround::eClassVar2D)
	{
		const double spawn_offset = -1;
		g->kind = 1;
		g->g2 = std::shared_ptr<Var2DProbe>(new Var2DProbe());
		tVector bound_min = tVector(-half_width + spawn_offset, 0, 0, 0);
		tVector bound_max = tVector(half_width + spawn_offset, 0, 0, 0);
		g->g2->SetTerrainFunc(cTerrainGen2D::GetTerrainFunc(g->params.mType));
		g->g2->Init(g->world, g->params, bound_min, bound_max);
	}
	else if (ground_class == cGround::eClassVar3D)
	{
		const double spawn_offset = 0;
		g->kind = 2;
		g->g3 = std::shared_ptr<Var3DProbe>(new Var3DProbe());
		tVector bound_min = tVector(-half_width + spawn_offset, 0, -half_width + spawn_offset, 0);
		tVector bound_max = tVector(half_width + spawn_offset, 0, half_width + spawn_offset, 0);
		g->g3->SetTerrainFunc(cTerrainGen3D::GetTerrainFunc(g->params.mType));
		g->g3->Init(g->world, g->params, bound_min, bound_max);
	}
	else
	{
		delete g;
		return nullptr;
	}
	return g;
}
void trlr_ground_free(void* h) { delete static_cast<Ground*>(h); }
int trlr_ground_kind(void* h) { return static_cast<Ground*>(h)->kind; }
int trlr_ground_num_pieces(void* h)
{
	Ground* g = static_cast<Ground*>(h);
	return g->kind == 1 ? g->g2->GetNumSegments() : g->g3->GetNumSlabs();
}
// cGround::Update(time, bound_min, bound_max): segment / slab scrolling
void trlr_ground_update(void* h, double time_elapsed, const double* bound_min, const double* bound_max)
{
	Ground* g = static_cast<Ground*>(h);
	tVector bmin(bound_min[0], bound_min[1], bound_min[2], 0), bmax(bound_max[0], bound_max[1], bound_max[2], 0);
	g->base()->Update(time_elapsed, bmin, bmax);
}
// piece s in the reference's scan order (GetSegment(s) / GetSlab(s)): res[2] (x, z), origin[3] = GetPos(), scale[3] =
// GetScaling(), bounds[4] = min_x, max_x, min_z, max_z as ContainsPos tests them; returns the number of floats
int trlr_ground_piece_info(void* h, int s, int* res, double* origin, double* scale, double* bounds)
{
	Ground* g = static_cast<Ground*>(h);
	const double inf = std::numeric_limits<double>::infinity();
	if (g->kind == 1)
	{
		const cGroundVar2D::tSegment& seg = g->g2->Seg(s);
		res[0] = seg.GetGridWidth();
		res[1] = seg.GetGridLength();
		from_vec(seg.GetPos(), origin, 3);
		from_vec(seg.GetScaling(), scale, 3);
		bounds[0] = seg.GetMinX();
		bounds[1] = seg.GetMaxX();
		bounds[2] = -inf;
		bounds[3] = inf;
		return static_cast<int>(seg.mData.size());
	}
	const cGroundVar3D::tSlab& slab = g->g3->Slab(s);
	res[0] = slab.mResX;
	res[1] = slab.mResZ;
	from_vec(slab.GetPos(), origin, 3);
	from_vec(slab.GetScaling(), scale, 3);
	bounds[0] = slab.mMin[0];
	bounds[1] = slab.mMax[0];
	bounds[2] = slab.mMin[2];
	bounds[3] = slab.mMax[2];
	return static_cast<int>(slab.mData.size());
}
void trlr_ground_piece_data(void* h, int s, float* out)
{
	Ground* g = static_cast<Ground*>(h);
	const std::vector<float>& d = (g->kind == 1) ? g->g2->Seg(s).mData : g->g3->Slab(s).mData;
	std::memcpy(out, d.data(), d.size() * sizeof(float));
}
// cGround::SampleHeight(pos, valid) for `count` positions [count][3]
void trlr_ground_sample_height(void* h, int count, const double* pos, double* out_h, int* out_valid)
{
	Ground* g = static_cast<Ground*>(h);
	for (int i = 0; i < count; ++i)
	{
		bool valid = false;
		out_h[i] = g->base()->SampleHeight(tVector(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 0), valid);
		out_valid[i] = valid ? 1 : 0;
	}
}
// cTerrainGen2D::<Build*>(width, params, rand, out_data) alone, seeded cRand; type_str as in terrain files ("var2d_mixed");
// params [cTerrainGen2D::eParamsMax] or NULL for the defaults; returns the number of vertices written (<= cap)
int trlr_terrain_gen_2d(const char* type_str, double width, const double* params, unsigned long seed, float* out, int cap)
{
	cGround::eType type = cGround::eTypeVar2DFlat;
	cGround::ParseType(type_str, type);
	Eigen::VectorXd p;
	cTerrainGen2D::GetDefaultParams(p);
	if (params) p = to_vec(params, static_cast<int>(p.size()));
	cRand rand(seed);
	std::vector<float> data;
	(*cTerrainGen2D::GetTerrainFunc(type))(width, p, rand, data);
	int n = static_cast<int>(data.size());
	for (int i = 0; i < n && i < cap; ++i) out[i] = data[i];
	return n;
}
int trlr_terrain_gen_2d_num_params(void) { return cTerrainGen2D::eParamsMax; }
// cTerrainGen2D::LoadParams from the i-th "Params" entry of a terrain file
int trlr_terrain_params_2d(const char* param_file, int idx, double* out)
{
	cGround::tParams params;
	if (!parse_ground_params(param_file, params)) return -1;
	if (idx >= params.mParamArr.rows()) return -1;
	for (int i = 0; i < params.mParamArr.cols(); ++i) out[i] = params.mParamArr(idx, i);
	return static_cast<int>(params.mParamArr.cols());
}

}  // extern "C"
