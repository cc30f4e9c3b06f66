// Stand-in for the reference's sim/SimCharacter.h -- TEST INFRASTRUCTURE (oracle/_ref build only).
//
// The real cSimCharacter owns Bullet bodies.  The PD controllers read only the generalized state
// (GetPose / GetVel = cSimCharacter::BuildPose / BuildVel, sim/SimCharacter.cpp:1037-1115), the
// joint table and the body table, and hand their torques to ApplyControlForces.  This stand-in
// holds those as plain data set by the test wrapper (oracle/ref_wrap.cpp).
#pragma once
#include <vector>
#include "anim/KinTree.h"
#include "sim/Joint.h"

class cSimCharacter
{
public:
	cSimCharacter() {}
	void InitStub(const Eigen::MatrixXd& joint_mat, const Eigen::MatrixXd& body_defs)
	{
		mJointMat = joint_mat;
		mBodyDefs = body_defs;
		int num_joints = GetNumJoints();
		mJoints.resize(num_joints);
		for (int j = 0; j < num_joints; ++j)
		{
			mJoints[j].Bind(this, j);
		}
		mPose = Eigen::VectorXd::Zero(GetNumDof());
		mVel = Eigen::VectorXd::Zero(GetNumDof());
		mAppliedTau = Eigen::VectorXd::Zero(GetNumDof());
	}
	void SetPose(const Eigen::VectorXd& pose) { mPose = pose; }
	void SetVel(const Eigen::VectorXd& vel) { mVel = vel; }

	int GetNumJoints() const { return cKinTree::GetNumJoints(mJointMat); }
	int GetNumDof() const { return cKinTree::GetNumDof(mJointMat); }
	int GetParamOffset(int joint_id) const { return cKinTree::GetParamOffset(mJointMat, joint_id); }
	int GetParamSize(int joint_id) const { return cKinTree::GetParamSize(mJointMat, joint_id); }
	const Eigen::MatrixXd& GetJointMat() const { return mJointMat; }
	const Eigen::MatrixXd& GetBodyDefs() const { return mBodyDefs; }
	const Eigen::VectorXd& GetPose() const { return mPose; }
	const Eigen::VectorXd& GetVel() const { return mVel; }
	cJoint& GetJoint(int joint_id) { return mJoints[joint_id]; }
	const cJoint& GetJoint(int joint_id) const { return mJoints[joint_id]; }
	void ApplyControlForces(const Eigen::VectorXd& tau) { mAppliedTau = tau; }
	const Eigen::VectorXd& GetAppliedTau() const { return mAppliedTau; }

protected:
	Eigen::MatrixXd mJointMat;
	Eigen::MatrixXd mBodyDefs;
	Eigen::VectorXd mPose;
	Eigen::VectorXd mVel;
	Eigen::VectorXd mAppliedTau;
	std::vector<cJoint> mJoints;
};

// a joint is valid when it has a Bullet constraint: a non-root joint whose body exists
// (cSimCharacter::BuildJoints, reference sim/SimCharacter.cpp:843-905)
inline bool cJoint::IsValid() const
{
	const Eigen::MatrixXd& joint_mat = mChar->GetJointMat();
	return cKinTree::HasParent(joint_mat, mID) && cKinTree::IsValidBody(mChar->GetBodyDefs(), mID);
}
inline cKinTree::eJointType cJoint::GetType() const { return cKinTree::GetJointType(mChar->GetJointMat(), mID); }
inline double cJoint::GetTorqueLimit() const { return cKinTree::GetTorqueLimit(mChar->GetJointMat(), mID); }
inline double cJoint::GetForceLimit() const { return cKinTree::GetForceLimit(mChar->GetJointMat(), mID); }
inline void cJoint::BuildPose(Eigen::VectorXd& out_pose) const
{
	out_pose = mChar->GetPose().segment(mChar->GetParamOffset(mID), mChar->GetParamSize(mID));
}
inline void cJoint::BuildVel(Eigen::VectorXd& out_vel) const
{
	out_vel = mChar->GetVel().segment(mChar->GetParamOffset(mID), mChar->GetParamSize(mID));
}
inline void cJoint::CalcRotation(tVector& out_axis, double& out_theta) const
{
	// reference sim/Joint.cpp:70-93,565-569,624-628
	out_axis = tVector(0, 0, 1, 0);
	out_theta = 0;
	switch (GetType())
	{
	case cKinTree::eJointTypeRevolute:
		out_theta = mChar->GetPose()[mChar->GetParamOffset(mID)];
		break;
	case cKinTree::eJointTypeSpherical:
	{
		Eigen::VectorXd pose;
		BuildPose(pose);
		cMathUtil::QuaternionToAxisAngle(cMathUtil::VecToQuat(pose), out_axis, out_theta);
		break;
	}
	default:
		break;
	}
}
inline tMatrix cJoint::BuildWorldTrans() const
{
	return cKinTree::JointWorldTrans(mChar->GetJointMat(), mChar->GetPose(), mID);
}
inline tQuaternion cJoint::CalcWorldRotation() const
{
	return cMathUtil::RotMatToQuaternion(BuildWorldTrans());
}
inline void cJoint::CalcWorldRotation(tVector& out_axis, double& out_theta) const
{
	cMathUtil::RotMatToAxisAngle(BuildWorldTrans(), out_axis, out_theta);
}
inline tVector cJoint::CalcAxisWorld() const
{
	// reference sim/Joint.cpp:46-53: the joint's local z axis in world coordinates
	tMatrix trans = BuildWorldTrans();
	tVector axis = trans * tVector(0, 0, 1, 0);
	return axis;
}
inline double cJoint::CalcDisplacementPrismatic() const { return mChar->GetPose()[mChar->GetParamOffset(mID)]; }
inline double cJoint::CalcDisplacementVelPrismatic() const { return mChar->GetVel()[mChar->GetParamOffset(mID)]; }
