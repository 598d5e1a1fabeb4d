// Minimal stand-in for the subset of Sophus (SO3d) the kinematic-icp API surface uses.  Present ONLY because Sophus
// is not installed in the build image; formulas follow Sophus' so3.hpp (quaternion storage, the
// SO3(quaternion) constructor normalises, exp/log with the small-angle Taylor branches at epsilon = 1e-10).
#pragma once
#include <Eigen/Core>
#include <cmath>

#define KICP_COMPAT_SOPHUS 1

namespace Sophus {

template <class Scalar>
struct Constants {
    static Scalar epsilon() { return Scalar(1e-10); }
    static Scalar pi() { return Scalar(3.141592653589793238462643383279502884); }
};

template <class Scalar_>
class SO3 {
public:
    using Scalar = Scalar_;
    using Point = Eigen::Matrix<Scalar, 3, 1>;
    using Tangent = Eigen::Matrix<Scalar, 3, 1>;
    using Transformation = Eigen::Matrix<Scalar, 3, 3>;
    using QuaternionType = Eigen::Quaternion<Scalar>;
    struct TangentAndTheta {
        Tangent tangent;
        Scalar theta;
    };

    SO3() : unit_quaternion_(Scalar(1), Scalar(0), Scalar(0), Scalar(0)) {}
    explicit SO3(const QuaternionType &quat) : unit_quaternion_(quat) { normalize(); }
    explicit SO3(const Transformation &R) {  // rotation matrix -> quaternion (Eigen's algorithm)
        const Scalar t = R.trace();
        Scalar w, x, y, z;
        if (t > Scalar(0)) {
            Scalar s = std::sqrt(t + Scalar(1));
            w = Scalar(0.5) * s;
            s = Scalar(0.5) / s;
            x = (R(2, 1) - R(1, 2)) * s, y = (R(0, 2) - R(2, 0)) * s, z = (R(1, 0) - R(0, 1)) * s;
        } else {
            int i = 0;
            if (R(1, 1) > R(0, 0)) i = 1;
            if (R(2, 2) > R(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            Scalar s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + Scalar(1));
            Scalar q[3];
            q[i] = Scalar(0.5) * s;
            s = Scalar(0.5) / s;
            w = (R(k, j) - R(j, k)) * s;
            q[j] = (R(j, i) + R(i, j)) * s;
            q[k] = (R(k, i) + R(i, k)) * s;
            x = q[0], y = q[1], z = q[2];
        }
        unit_quaternion_ = QuaternionType(w, x, y, z);
        normalize();
    }

    const QuaternionType &unit_quaternion() const { return unit_quaternion_; }
    Transformation matrix() const { return unit_quaternion_.toRotationMatrix(); }
    SO3 inverse() const {
        SO3 r;
        r.unit_quaternion_ = unit_quaternion_.conjugate();
        return r;
    }
    SO3 operator*(const SO3 &other) const { return SO3(unit_quaternion_ * other.unit_quaternion_); }
    // uv = 2 (q.vec x p); p + w uv + q.vec x uv
    Point operator*(const Point &p) const { return unit_quaternion_._transformVector(p); }

    static Transformation hat(const Tangent &omega) {
        Transformation Omega;
        Omega(0, 1) = -omega(2), Omega(0, 2) = omega(1);
        Omega(1, 0) = omega(2), Omega(1, 2) = -omega(0);
        Omega(2, 0) = -omega(1), Omega(2, 1) = omega(0);
        return Omega;
    }
    static SO3 exp(const Tangent &omega) {
        Scalar theta;
        return expAndTheta(omega, &theta);
    }
    static SO3 expAndTheta(const Tangent &omega, Scalar *theta) {
        const Scalar theta_sq = omega.squaredNorm();
        Scalar imag_factor, real_factor;
        if (theta_sq < Constants<Scalar>::epsilon() * Constants<Scalar>::epsilon()) {
            *theta = Scalar(0);
            const Scalar theta_po4 = theta_sq * theta_sq;
            imag_factor = Scalar(0.5) - Scalar(1.0 / 48.0) * theta_sq + Scalar(1.0 / 3840.0) * theta_po4;
            real_factor = Scalar(1) - Scalar(1.0 / 8.0) * theta_sq + Scalar(1.0 / 384.0) * theta_po4;
        } else {
            *theta = std::sqrt(theta_sq);
            const Scalar half_theta = Scalar(0.5) * (*theta);
            imag_factor = std::sin(half_theta) / (*theta);
            real_factor = std::cos(half_theta);
        }
        SO3 q;
        q.unit_quaternion_ = QuaternionType(real_factor, imag_factor * omega.x(), imag_factor * omega.y(), imag_factor * omega.z());
        return q;
    }
    Tangent log() const { return logAndTheta().tangent; }
    TangentAndTheta logAndTheta() const {
        TangentAndTheta J;
        const Scalar squared_n = unit_quaternion_.vec().squaredNorm();
        const Scalar w = unit_quaternion_.w();
        Scalar two_atan_nbyw_by_n;
        if (squared_n < Constants<Scalar>::epsilon() * Constants<Scalar>::epsilon()) {
            const Scalar squared_w = w * w;
            two_atan_nbyw_by_n = Scalar(2) / w - Scalar(2.0 / 3.0) * (squared_n) / (w * squared_w);
            J.theta = Scalar(2) * squared_n / w;
        } else {
            const Scalar n = std::sqrt(squared_n);
            const Scalar atan_nbyw = (w < Scalar(0)) ? Scalar(std::atan2(-n, -w)) : Scalar(std::atan2(n, w));
            two_atan_nbyw_by_n = Scalar(2) * atan_nbyw / n;
            J.theta = two_atan_nbyw_by_n * n;
        }
        J.tangent = unit_quaternion_.vec() * two_atan_nbyw_by_n;
        return J;
    }

private:
    void normalize() {
        const Scalar length = unit_quaternion_.norm();
        unit_quaternion_.coeffs() /= length;
    }
    QuaternionType unit_quaternion_;
};

using SO3d = SO3<double>;

}  // namespace Sophus
