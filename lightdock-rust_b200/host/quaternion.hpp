// quaternion.hpp — host-side Quaternion, mirroring the public surface of src/qt.rs.
// `rotate` exists for API parity and the unit tests; on the hot path the rotation runs on the device
// (csrc/ld_kernels.cuh, transform_kernel).  `slerp` is the host-side move operator of the GSO.
#pragma once
#include <cmath>
#include <limits>

namespace lightdock {

constexpr double LINEAR_THRESHOLD = 0.9995;  // src/constants.rs:11

struct Quaternion {
  double w = 1.0, x = 0.0, y = 0.0, z = 0.0;  // Default: identity (src/qt.rs:106-115)
  Quaternion() = default;
  Quaternion(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}

  Quaternion conjugate() const { return {w, -x, -y, -z}; }
  double dot(const Quaternion &o) const { return w * o.w + x * o.x + y * o.y + z * o.z; }
  double norm2() const { return w * w + x * x + y * y + z * z; }
  double norm() const { return std::sqrt(w * w + x * x + y * y + z * z); }
  void normalize() {
    const double n = norm();
    w /= n; x /= n; y /= n; z /= n;
  }
  Quaternion operator-() const { return {-w, -x, -y, -z}; }
  Quaternion operator+(const Quaternion &o) const { return {w + o.w, x + o.x, y + o.y, z + o.z}; }
  Quaternion operator-(const Quaternion &o) const { return {w - o.w, x - o.x, y - o.y, z - o.z}; }
  Quaternion operator*(double s) const { return {s * w, s * x, s * y, s * z}; }
  Quaternion operator/(double s) const { return {w / s, x / s, y / s, z / s}; }
  Quaternion operator*(const Quaternion &o) const {  // src/qt.rs:174-185
    return {w * o.w - x * o.x - y * o.y - z * o.z, w * o.x + x * o.w + y * o.z - z * o.y,
            w * o.y - x * o.z + y * o.w + z * o.x, w * o.z + x * o.y - y * o.x + z * o.w};
  }
  Quaternion inverse() const { return conjugate() / norm2(); }
  double distance(const Quaternion &o) const {
    const double d = dot(o);
    return 1.0 - d * d;
  }
  // q * (0,v) * q^-1, src/qt.rs:57-61
  void rotate(const double v[3], double out[3]) const {
    const Quaternion r = (*this) * Quaternion(0.0, v[0], v[1], v[2]) * inverse();
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
  }
  Quaternion lerp(const Quaternion &o, double t) const { return (*this) * (1.0 - t) + o * t; }
  // src/qt.rs:67-91
  Quaternion slerp(const Quaternion &other, double t) const {
    Quaternion q1 = *this, q2 = other;
    q1.normalize();
    q2.normalize();
    double q_dot = q1.dot(q2);
    if (q_dot < 0.0) {  // avoid the long path
      q1 = -q1;
      q_dot *= -1.0;
    }
    if (q_dot > LINEAR_THRESHOLD) {
      Quaternion result = q1 + (q2 - q1) * t;
      result.normalize();
      return result;
    }
    q_dot = std::fmax(std::fmin(q_dot, 1.0), -1.0);
    const double omega = std::acos(q_dot);
    const double so = std::sin(omega);
    return q1 * (std::sin((1.0 - t) * omega) / so) + q2 * (std::sin(t * omega) / so);
  }
  // PartialEq of the reference: |delta| < f64::EPSILON per component (src/qt.rs:7-9,143-150)
  bool operator==(const Quaternion &o) const {
    const double eps = std::numeric_limits<double>::epsilon();
    return std::fabs(w - o.w) < eps && std::fabs(x - o.x) < eps && std::fabs(y - o.y) < eps &&
           std::fabs(z - o.z) < eps;
  }
};

}  // namespace lightdock
