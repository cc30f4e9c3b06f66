// Stand-in for the slice of Bullet the reference's terrain classes touch -- TEST INFRASTRUCTURE
// (oracle/_ref build only).  Bullet is an un-vendored dependency of the reference;
// sim/GroundVar2D.cpp / GroundVar3D.cpp (compiled UNMODIFIED against this header) use it only to
// store a heightfield, its local scaling and a float world transform, and to read back an AABB.
// Restated from Bullet's published behaviour (btScalar = float, no BT_USE_DOUBLE_PRECISION in the
// reference's premake4.lua):
//   * btHeightfieldTerrainShape::getAabb: half extents = (localAabbMax - localAabbMin) * localScaling * 0.5
//     around the transform origin, plus the concave-shape margin, which is 0;
//   * local AABB for up axis 1: (0, minHeight, 0) .. (width-1, maxHeight, length-1).
#pragma once
#include <cmath>

typedef float btScalar;
enum PHY_ScalarType { PHY_FLOAT, PHY_DOUBLE, PHY_INTEGER, PHY_SHORT, PHY_FIXEDPOINT88, PHY_UCHAR };

class btVector3
{
public:
	btScalar m_floats[4];
	btVector3() { m_floats[0] = m_floats[1] = m_floats[2] = m_floats[3] = 0; }
	btVector3(btScalar x, btScalar y, btScalar z) { m_floats[0] = x; m_floats[1] = y; m_floats[2] = z; m_floats[3] = 0; }
	btScalar& operator[](int i) { return m_floats[i]; }
	const btScalar& operator[](int i) const { return m_floats[i]; }
	btVector3 operator+(const btVector3& o) const { return btVector3(m_floats[0] + o[0], m_floats[1] + o[1], m_floats[2] + o[2]); }
	btVector3 operator-(const btVector3& o) const { return btVector3(m_floats[0] - o[0], m_floats[1] - o[1], m_floats[2] - o[2]); }
	btVector3 operator*(const btVector3& o) const { return btVector3(m_floats[0] * o[0], m_floats[1] * o[1], m_floats[2] * o[2]); }
	btVector3 operator*(btScalar s) const { return btVector3(m_floats[0] * s, m_floats[1] * s, m_floats[2] * s); }
};
inline btVector3 operator*(btScalar s, const btVector3& v) { return v * s; }

class btTransform
{
	btVector3 m_origin;  // the terrain pieces are never rotated: basis = identity

public:
	const btVector3& getOrigin() const { return m_origin; }
	void setOrigin(const btVector3& o) { m_origin = o; }
};

class btMotionState
{
public:
	virtual ~btMotionState() {}
	virtual void setWorldTransform(const btTransform&) {}
};
class btDefaultMotionState : public btMotionState
{
};

class btCollisionShape
{
protected:
	btVector3 m_localScaling;

public:
	btCollisionShape() : m_localScaling(1, 1, 1) {}
	virtual ~btCollisionShape() {}
	virtual void setLocalScaling(const btVector3& s) { m_localScaling = s; }
	virtual const btVector3& getLocalScaling() const { return m_localScaling; }
	virtual void getAabb(const btTransform& t, btVector3& aabb_min, btVector3& aabb_max) const = 0;
};

class btHeightfieldTerrainShape : public btCollisionShape
{
	btVector3 m_localAabbMin, m_localAabbMax;

public:
	btHeightfieldTerrainShape(int stick_width, int stick_length, const void*, btScalar, btScalar min_height,
							  btScalar max_height, int up_axis, PHY_ScalarType, bool)
	{
		(void)up_axis;  // the reference always passes 1 (y up)
		m_localAabbMin = btVector3(0, min_height, 0);
		m_localAabbMax = btVector3(btScalar(stick_width - 1), max_height, btScalar(stick_length - 1));
	}
	virtual void getAabb(const btTransform& t, btVector3& aabb_min, btVector3& aabb_max) const
	{
		btVector3 half_extents = (m_localAabbMax - m_localAabbMin) * m_localScaling * btScalar(0.5);
		btVector3 center = t.getOrigin();
		btVector3 extent = half_extents;  // identity basis; margin of a concave shape = 0
		aabb_min = center - extent;
		aabb_max = center + extent;
	}
};

class btRigidBody
{
	btTransform m_worldTransform;
	btCollisionShape* m_shape;
	btScalar m_friction;
	void* m_user;

public:
	struct btRigidBodyConstructionInfo
	{
		btCollisionShape* m_collisionShape;
		btRigidBodyConstructionInfo(btScalar, btMotionState*, btCollisionShape* shape, const btVector3&) : m_collisionShape(shape) {}
	};
	btRigidBody(const btRigidBodyConstructionInfo& info) : m_shape(info.m_collisionShape), m_friction(0), m_user(0) {}
	const btTransform& getWorldTransform() const { return m_worldTransform; }
	void setWorldTransform(const btTransform& t) { m_worldTransform = t; }
	void getAabb(btVector3& aabb_min, btVector3& aabb_max) const { m_shape->getAabb(m_worldTransform, aabb_min, aabb_max); }
	void setFriction(btScalar f) { m_friction = f; }
	void setUserPointer(void* p) { m_user = p; }
	bool isKinematicObject() const { return false; }
	btMotionState* getMotionState() { return 0; }
};
