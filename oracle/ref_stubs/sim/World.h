// Stand-in for the reference's sim/World.h -- TEST INFRASTRUCTURE (oracle/_ref build only).
// The terrain classes need from cWorld: the world scale, and position / AABB access to a static
// body.  Those follow the reference's sim/World.cpp:312-362,498-508 (positions pass through
// Bullet's float transform: SURVEY.md Q9).
#pragma once
#include <memory>
#include "BulletCollision/CollisionShapes/btHeightfieldTerrainShape.h"
#include "anim/KinTree.h"
#include "util/MathUtil.h"

class cSimObj;

class cWorld
{
public:
	enum eContactFlag
	{
		eContactFlagCharacter = 0x1,
		eContactFlagEnvironment = 0x1 << 1,
		eContactFlagObject = 0x1 << 2,
		eContactFlagAll = -1
	};
	struct tConstraintHandle
	{
		bool IsValid() const { return false; }
	};

	explicit cWorld(double scale) : mScale(scale) {}
	double GetScale() const { return mScale; }
	tVector GetPos(const cSimObj* obj) const;
	void SetPos(const tVector& pos, cSimObj* out_obj) const;
	void CalcAABB(const cSimObj* obj, tVector& out_min, tVector& out_max) const;

protected:
	double mScale;
};
