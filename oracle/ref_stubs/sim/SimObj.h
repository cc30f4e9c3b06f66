// Stand-in for the reference's sim/SimObj.h -- TEST INFRASTRUCTURE (oracle/_ref build only).
// cGround and the terrain pieces derive from cSimObj; they use Init / RemoveFromWorld, GetPos /
// SetPos, CalcAABB and UpdateContact.  Everything else of the Bullet-backed original is absent.
#pragma once
#include <memory>
#include "sim/World.h"

class cSimObj : public btDefaultMotionState
{
public:
	enum eType
	{
		eTypeDynamic,
		eTypeStatic,
		eTypeMax
	};

	virtual ~cSimObj() {}

	virtual tVector GetPos() const { return mWorld->GetPos(this); }
	virtual void SetPos(const tVector& pos) { mWorld->SetPos(pos, this); }
	virtual void CalcAABB(tVector& out_min, tVector& out_max) const { mWorld->CalcAABB(this, out_min, out_max); }
	virtual void UpdateContact(int, int) {}
	virtual bool HasSimBody() const { return mSimBody != nullptr; }
	virtual const std::unique_ptr<btRigidBody>& GetSimBody() const { return mSimBody; }
	virtual const std::shared_ptr<cWorld>& GetWorld() const { return mWorld; }

protected:
	std::shared_ptr<cWorld> mWorld;
	std::unique_ptr<btRigidBody> mSimBody;
	std::unique_ptr<btCollisionShape> mShape;
	eType mType;

	cSimObj() : mType(eTypeDynamic) {}

	virtual void Init(const std::shared_ptr<cWorld>& world)
	{
		// reference sim/SimObj.cpp:271-287
		if (mSimBody != nullptr)
		{
			mSimBody->setUserPointer(this);
		}
		mWorld = world;
	}
	virtual void RemoveFromWorld()
	{
		// reference sim/SimObj.cpp:289-309
		if (mWorld != nullptr && mSimBody != nullptr)
		{
			mWorld.reset();
			mSimBody.reset();
		}
	}
};

// reference sim/World.cpp:333-341
inline tVector cWorld::GetPos(const cSimObj* obj) const
{
	auto& body = obj->GetSimBody();
	btTransform trans = body->getWorldTransform();
	btVector3 origin = trans.getOrigin();
	tVector pos = tVector(origin[0], origin[1], origin[2], 0);
	pos /= GetScale();
	return pos;
}
// reference sim/World.cpp:343-362
inline void cWorld::SetPos(const tVector& pos, cSimObj* out_obj) const
{
	auto& body = out_obj->GetSimBody();
	btTransform trans = body->getWorldTransform();
	btScalar scale = static_cast<btScalar>(GetScale());
	trans.setOrigin(scale * btVector3(static_cast<btScalar>(pos[0]), static_cast<btScalar>(pos[1]), static_cast<btScalar>(pos[2])));
	body->setWorldTransform(trans);
}
// reference sim/World.cpp:498-508
inline void cWorld::CalcAABB(const cSimObj* obj, tVector& out_min, tVector& out_max) const
{
	auto& body = obj->GetSimBody();
	btVector3 bt_min;
	btVector3 bt_max;
	body->getAabb(bt_min, bt_max);
	double scale = GetScale();
	out_min = tVector(bt_min[0], bt_min[1], bt_min[2], 0) / scale;
	out_max = tVector(bt_max[0], bt_max[1], bt_max[2], 0) / scale;
}
