// Stand-in for the reference's sim/Joint.h -- TEST INFRASTRUCTURE (oracle/_ref build only).
//
// The real cJoint wraps a Bullet constraint.  The PD controllers (sim/PDController.cpp,
// sim/ExpPDController.cpp, sim/ImpPDController.cpp -- compiled UNMODIFIED against this header)
// only read a joint's generalized coordinates through it, so this stand-in serves them from
// the character's pose / vel vectors, i.e. exactly the data contract of cJoint::BuildPose /
// BuildVel (reference sim/Joint.cpp:353-417,446-534): revolute = angle about local z,
// prismatic = displacement along local x, spherical = quaternion (w,x,y,z) / (omega, 0).
// World rotations come from cKinTree::JointWorldTrans of the same pose (= Bullet's body
// transform times mJointChildTrans in a consistent simulation, sim/Joint.cpp:126-131).
#pragma once
#include "anim/KinTree.h"
#include "sim/SpAlg.h"

class cSimCharacter;

class cJoint
{
public:
	cJoint() : mChar(nullptr), mID(gInvalidIdx) {}
	void Bind(const cSimCharacter* character, int id) { mChar = character; mID = id; }

	bool IsValid() const;
	cKinTree::eJointType GetType() const;
	double GetTorqueLimit() const;
	double GetForceLimit() const;

	void CalcRotation(tVector& out_axis, double& out_theta) const;
	tQuaternion CalcWorldRotation() const;
	void CalcWorldRotation(tVector& out_axis, double& out_theta) const;
	tVector CalcAxisWorld() const;
	tMatrix BuildWorldTrans() const;

	void BuildPose(Eigen::VectorXd& out_pose) const;
	void BuildVel(Eigen::VectorXd& out_vel) const;
	double CalcDisplacementPrismatic() const;
	double CalcDisplacementVelPrismatic() const;

protected:
	const cSimCharacter* mChar;
	int mID;
};
