// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for ethz-asl/minkindr
// (kindr::minimal), a third-party dependency of the reference pulled at
// unpinned git HEAD by voxblox_https.rosinstall and absent from this image.
// Restates the published behaviour the hot path relies on:
//   RotationQuaternionTemplate::rotate(v)    = q * v   (Eigen _transformVector)
//   QuatTransformationTemplate::operator*(v) = q.rotate(v) + t
//   inverse()                                = (q^-1, -(q^-1 * t))
// and, for the ICP path (src/alignment/icp.cc:170-216), restated from minkindr's published
// rotation-quaternion-inl.h / quat-transformation-inl.h (the library itself is not in the image, so
// these cannot be pinned bit for bit -- the ICP parity bound is a float tolerance):
//   T1 * T2       = (q1 * q2, t1 + q1.rotate(t2))
//   T.log()       = [t ; q.log()]                       (6-vector, translation first)
//   exp([t ; w])  = (RotationQuaternion::exp(w), t)
//   q.log()       = 2 * scale * imag(q), scale = acos(eta)/|a| when |eta| < |a| (sign of eta kept so
//                   that log(-q) = log(q)), else +-asin(|a|)/|a| with the series 1 + x^2/6 near 0
//   exp(w)        = (cos(theta/2), w * sin(theta/2)/theta), theta = |w|, series 1/2 + theta^2/48 near 0
//                   (Grassia 1998), evaluated in double as minkindr does
//   isValidRotationMatrix(R): |det R - 1| and max |R R^T - I| below a threshold
#ifndef VBX_ORACLE_SHIM_KINDR_QUAT_TRANSFORMATION_H_
#define VBX_ORACLE_SHIM_KINDR_QUAT_TRANSFORMATION_H_

#include <Eigen/Core>

#include <cmath>
#include <limits>

namespace kindr {
namespace minimal {

template <typename Scalar>
class RotationQuaternionTemplate {
 public:
  typedef Eigen::Quaternion<Scalar> Implementation;
  typedef Eigen::Matrix<Scalar, 3, 1> Vector3;

  typedef Eigen::Matrix<Scalar, 3, 3> RotationMatrix;

  RotationQuaternionTemplate() {}
  RotationQuaternionTemplate(Scalar w, Scalar x, Scalar y, Scalar z) : q_(w, x, y, z) {}
  explicit RotationQuaternionTemplate(const Implementation& q) : q_(q) {}
  explicit RotationQuaternionTemplate(const RotationMatrix& matrix) : q_(matrix) {}

  RotationQuaternionTemplate operator*(const RotationQuaternionTemplate& rhs) const {
    return RotationQuaternionTemplate(q_ * rhs.q_);
  }

  static bool isValidRotationMatrix(const RotationMatrix& matrix) {
    // minkindr's default threshold is not recoverable here; a float-friendly one is used.  Rotations
    // built from an SVD pass with a wide margin and non-finite input is caught by the caller
    // (icp.h:183-186), so the exact value does not influence the results
    const Scalar threshold = static_cast<Scalar>(1000) * std::numeric_limits<Scalar>::epsilon();
    if (std::fabs(matrix.determinant() - static_cast<Scalar>(1.0)) > threshold) return false;
    const RotationMatrix e = matrix * matrix.transpose() - RotationMatrix::Identity();
    if (e.cwiseAbs().maxCoeff() > threshold) return false;
    return true;
  }

  static RotationQuaternionTemplate exp(const Vector3& dx) {
    const double x = dx[0], y = dx[1], z = dx[2];
    const double theta = std::sqrt(x * x + y * y + z * z);
    double na;
    if (theta < std::pow(std::numeric_limits<double>::epsilon(), 0.25)) {
      na = 0.5 + (theta * theta) * (1.0 / 48.0);
    } else {
      na = std::sin(theta * 0.5) / theta;
    }
    const double ct = std::cos(theta * 0.5);
    return RotationQuaternionTemplate(static_cast<Scalar>(ct), static_cast<Scalar>(x * na),
                                      static_cast<Scalar>(y * na), static_cast<Scalar>(z * na));
  }
  Vector3 log() const {
    const Vector3 a = q_.vec();
    const Scalar na = a.norm();
    const Scalar eta = q_.w();
    Scalar scale;
    if (std::fabs(eta) < na) {
      scale = eta >= Scalar(0) ? std::acos(eta) / na : -std::acos(-eta) / na;
    } else {
      const Scalar s = arcSinXOverX(na);
      scale = eta > Scalar(0) ? s : -s;
    }
    return a * (Scalar(2.0) * scale);
  }

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
  static Scalar arcSinXOverX(Scalar x) {
    if (std::fabs(x) < std::pow(std::numeric_limits<Scalar>::epsilon(), Scalar(0.25))) {
      return Scalar(1.0) + x * x * Scalar(1.0 / 6.0);
    }
    return std::asin(x) / x;
  }
  Implementation q_;
};

template <typename Scalar>
class QuatTransformationTemplate {
 public:
  typedef Eigen::Matrix<Scalar, 3, 1> Vector3;
  typedef Vector3 Position;
  typedef RotationQuaternionTemplate<Scalar> Rotation;
  typedef Eigen::Matrix<Scalar, 6, 1> Vector6;

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
  QuatTransformationTemplate operator*(const QuatTransformationTemplate& rhs) const {
    return QuatTransformationTemplate(q_ * rhs.q_, t_ + q_.rotate(rhs.t_));
  }
  Vector6 log() const {
    const Vector3 w = q_.log();
    Vector6 v;
    for (int i = 0; i < 3; ++i) {
      v[i] = t_[i];
      v[3 + i] = w[i];
    }
    return v;
  }
  static QuatTransformationTemplate exp(const Vector6& v) {
    return QuatTransformationTemplate(Rotation::exp(Vector3(v[3], v[4], v[5])), Position(v[0], v[1], v[2]));
  }

 private:
  Rotation q_;
  Position t_;
};

}  // namespace minimal
}  // namespace kindr

#endif  // VBX_ORACLE_SHIM_KINDR_QUAT_TRANSFORMATION_H_
