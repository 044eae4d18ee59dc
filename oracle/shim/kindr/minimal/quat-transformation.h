// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for ethz-asl/minkindr
// (kindr::minimal), a third-party dependency of the reference pulled at
// unpinned git HEAD by voxblox_https.rosinstall and absent from this image.
// Restates the published behaviour the hot path relies on:
//   RotationQuaternionTemplate::rotate(v)    = q * v   (Eigen _transformVector)
//   QuatTransformationTemplate::operator*(v) = q.rotate(v) + t
//   inverse()                                = (q^-1, -(q^-1 * t))
#ifndef VBX_ORACLE_SHIM_KINDR_QUAT_TRANSFORMATION_H_
#define VBX_ORACLE_SHIM_KINDR_QUAT_TRANSFORMATION_H_

#include <Eigen/Core>

namespace kindr {
namespace minimal {

template <typename Scalar>
class RotationQuaternionTemplate {
 public:
  typedef Eigen::Quaternion<Scalar> Implementation;
  typedef Eigen::Matrix<Scalar, 3, 1> Vector3;

  RotationQuaternionTemplate() {}
  RotationQuaternionTemplate(Scalar w, Scalar x, Scalar y, Scalar z) : q_(w, x, y, z) {}
  explicit RotationQuaternionTemplate(const Implementation& q) : q_(q) {}

  Scalar w() const { return q_.w(); }
  Scalar x() const { return q_.x(); }
  Scalar y() const { return q_.y(); }
  Scalar z() const { return q_.z(); }
  const Implementation& toImplementation() const { return q_; }

  Vector3 rotate(const Vector3& v) const { return q_ * v; }
  Vector3 inverseRotate(const Vector3& v) const { return q_.conjugate() * v; }
  RotationQuaternionTemplate inverse() const {
    return RotationQuaternionTemplate(q_.conjugate());
  }

 private:
  Implementation q_;
};

template <typename Scalar>
class QuatTransformationTemplate {
 public:
  typedef Eigen::Matrix<Scalar, 3, 1> Vector3;
  typedef Vector3 Position;
  typedef RotationQuaternionTemplate<Scalar> Rotation;

  QuatTransformationTemplate() : t_(Position::Zero()) {}
  QuatTransformationTemplate(const Rotation& q, const Position& t) : q_(q), t_(t) {}
  QuatTransformationTemplate(const typename Rotation::Implementation& q,
                             const Position& t)
      : q_(q), t_(t) {}

  const Position& getPosition() const { return t_; }
  const Rotation& getRotation() const { return q_; }

  Vector3 transform(const Vector3& v) const { return q_.rotate(v) + t_; }
  Vector3 operator*(const Vector3& v) const { return transform(v); }
  QuatTransformationTemplate inverse() const {
    return QuatTransformationTemplate(q_.inverse(), -q_.inverseRotate(t_));
  }

 private:
  Rotation q_;
  Position t_;
};

}  // namespace minimal
}  // namespace kindr

#endif  // VBX_ORACLE_SHIM_KINDR_QUAT_TRANSFORMATION_H_
