// Minimal stand-in for the subset of Sophus (SE3d) the kinematic-icp API surface uses.  Present ONLY because Sophus
// is not installed in the build image; formulas follow Sophus' se3.hpp.
#pragma once
#include "so3.hpp"

namespace Sophus {

template <class Scalar_>
class SE3 {
public:
    using Scalar = Scalar_;
    using Point = Eigen::Matrix<Scalar, 3, 1>;
    using Tangent = Eigen::Matrix<Scalar, 6, 1>;  // (upsilon, omega)
    using Matrix3 = Eigen::Matrix<Scalar, 3, 3>;
    using SO3Type = SO3<Scalar>;
    using QuaternionType = Eigen::Quaternion<Scalar>;

    SE3() {}
    SE3(const SO3Type &so3, const Point &translation) : so3_(so3), translation_(translation) {}
    SE3(const QuaternionType &q, const Point &translation) : so3_(q), translation_(translation) {}
    SE3(const Matrix3 &R, const Point &translation) : so3_(R), translation_(translation) {}

    const SO3Type &so3() const { return so3_; }
    SO3Type &so3() { return so3_; }
    const Point &translation() const { return translation_; }
    Point &translation() { return translation_; }
    const QuaternionType &unit_quaternion() const { return so3_.unit_quaternion(); }
    Matrix3 rotationMatrix() const { return so3_.matrix(); }

    SE3 operator*(const SE3 &other) const { return SE3(so3_ * other.so3_, translation_ + so3_ * other.translation_); }
    SE3 &operator*=(const SE3 &other) {
        *this = *this * other;
        return *this;
    }
    Point operator*(const Point &p) const { return so3_ * p + translation_; }
    SE3 inverse() const {
        const SO3Type invR = so3_.inverse();
        return SE3(invR, invR * (translation_ * Scalar(-1)));
    }

    static SE3 exp(const Tangent &a) {
        const Point omega = a.template tail<3>();
        Scalar theta;
        const SO3Type so3 = SO3Type::expAndTheta(omega, &theta);
        const Matrix3 Omega = SO3Type::hat(omega);
        const Matrix3 Omega_sq = Omega * Omega;
        Matrix3 V;
        if (theta < Constants<Scalar>::epsilon()) {
            V = so3.matrix();
        } else {
            const Scalar theta_sq = theta * theta;
            V = (Matrix3::Identity() + Omega * ((Scalar(1) - std::cos(theta)) / (theta_sq)) +
                 Omega_sq * ((theta - std::sin(theta)) / (theta_sq * theta)));
        }
        const Point upsilon = a.template head<3>();
        return SE3(so3, V * upsilon);
    }
    Tangent log() const {
        Tangent upsilon_omega;
        const auto omega_and_theta = so3_.logAndTheta();
        const Scalar theta = omega_and_theta.theta;
        upsilon_omega.template tail<3>() = omega_and_theta.tangent;
        const Matrix3 Omega = SO3Type::hat(omega_and_theta.tangent);
        Matrix3 V_inv;
        if (std::abs(theta) < Constants<Scalar>::epsilon()) {
            V_inv = Matrix3::Identity() - Omega * Scalar(0.5) + (Omega * Omega) * Scalar(1. / 12.);
        } else {
            const Scalar half_theta = Scalar(0.5) * theta;
            V_inv = (Matrix3::Identity() - Omega * Scalar(0.5) +
                     (Omega * Omega) *
                         ((Scalar(1) - theta * std::cos(half_theta) / (Scalar(2) * std::sin(half_theta))) / (theta * theta)));
        }
        upsilon_omega.template head<3>() = V_inv * translation_;
        return upsilon_omega;
    }

private:
    SO3Type so3_;
    Point translation_;
};

using SE3d = SE3<double>;

}  // namespace Sophus
