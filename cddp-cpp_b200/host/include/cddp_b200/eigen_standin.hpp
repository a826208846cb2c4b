// Minimal stand-in for the subset of Eigen the host mirror uses (VectorXd / MatrixXd construction,
// Zero/Identity/Constant, size/rows/cols, element access, a few reductions).  The reference's public API is
// expressed in Eigen types (include/cddp-cpp/cddp_core/cddp_core.hpp); Eigen is not available in this image
// (SURVEY.md F1).  cddp.hpp includes the real <Eigen/Dense> when it exists and this file otherwise, and the
// mirror restricts itself to calls that mean the same thing in both.
#pragma once
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <stdexcept>
#include <vector>

namespace Eigen {

class VectorXd {
 public:
  VectorXd() = default;
  explicit VectorXd(long n) : v_((size_t)n, 0.0) {}
  VectorXd(std::initializer_list<double> l) : v_(l) {}
  static VectorXd Zero(long n) { return VectorXd(n); }
  static VectorXd Constant(long n, double c) {
    VectorXd r(n);
    for (auto &x : r.v_) x = c;
    return r;
  }
  long size() const { return (long)v_.size(); }
  double &operator()(long i) { return v_[(size_t)i]; }
  double operator()(long i) const { return v_[(size_t)i]; }
  double &operator[](long i) { return v_[(size_t)i]; }
  double operator[](long i) const { return v_[(size_t)i]; }
  double *data() { return v_.data(); }
  const double *data() const { return v_.data(); }
  bool isZero(double eps = 1e-12) const {
    for (double x : v_)
      if (std::fabs(x) > eps) return false;
    return true;
  }
  double norm() const {
    double s = 0;
    for (double x : v_) s += x * x;
    return std::sqrt(s);
  }
  VectorXd operator-(const VectorXd &o) const {
    VectorXd r(size());
    for (long i = 0; i < size(); ++i) r[i] = v_[(size_t)i] - o[i];
    return r;
  }
  VectorXd operator+(const VectorXd &o) const {
    VectorXd r(size());
    for (long i = 0; i < size(); ++i) r[i] = v_[(size_t)i] + o[i];
    return r;
  }
  VectorXd operator*(double c) const {
    VectorXd r(size());
    for (long i = 0; i < size(); ++i) r[i] = v_[(size_t)i] * c;
    return r;
  }
  double dot(const VectorXd &o) const {
    double s = 0;
    for (long i = 0; i < size(); ++i) s += v_[(size_t)i] * o[i];
    return s;
  }

 private:
  std::vector<double> v_;
};

class MatrixXd {  // storage order is an implementation detail: always address elements as (i, j)
 public:
  MatrixXd() = default;
  MatrixXd(long r, long c) : r_(r), c_(c), v_((size_t)(r * c), 0.0) {}
  static MatrixXd Zero(long r, long c) { return MatrixXd(r, c); }
  static MatrixXd Identity(long r, long c) {
    MatrixXd m(r, c);
    for (long i = 0; i < (r < c ? r : c); ++i) m(i, i) = 1.0;
    return m;
  }
  long rows() const { return r_; }
  long cols() const { return c_; }
  long size() const { return r_ * c_; }
  double &operator()(long i, long j) { return v_[(size_t)(i * c_ + j)]; }
  double operator()(long i, long j) const { return v_[(size_t)(i * c_ + j)]; }
  MatrixXd operator*(double s) const {
    MatrixXd m(*this);
    for (auto &x : m.v_) x *= s;
    return m;
  }
  VectorXd operator*(const VectorXd &x) const {
    VectorXd y(r_);
    for (long i = 0; i < r_; ++i) {
      double s = 0;
      for (long j = 0; j < c_; ++j) s += (*this)(i, j) * x[j];
      y[i] = s;
    }
    return y;
  }
  MatrixXd transpose() const {
    MatrixXd m(c_, r_);
    for (long i = 0; i < r_; ++i)
      for (long j = 0; j < c_; ++j) m(j, i) = (*this)(i, j);
    return m;
  }

 private:
  long r_ = 0, c_ = 0;
  std::vector<double> v_;
};

using Matrix3d = MatrixXd;
using VectorXi = std::vector<int>;

}  // namespace Eigen
