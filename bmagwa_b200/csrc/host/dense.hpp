// dense.hpp -- the k x k dense kit the host sampler needs (k = model size, tens of columns).
//
// Plays the role of the reference's Vector / SymmMatrix wrappers over BLAS, LAPACK and LINPACK
// (src/symmmatrix.cpp:144-265, src/dchex.f): upper-triangular Cholesky factor with O(k^2) append,
// delete (left circular shift + Givens re-triangularisation, what dchex job=2 does) and adjacent
// swap.  Stays on the host by design (north star: "the small Cholesky add/remove/exchange update
// stays on one GPU with the reference's RNG stream" -- it is O(k^2) flops on k ~ 20).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace bmg {

// Column-major square storage with a fixed leading dimension; only the upper triangle is meaningful.
class UpperMat {
 public:
  explicit UpperMat(int ld = 64) : ld_(ld), k_(0), a_((size_t)ld * ld, 0.0) {}
  int size() const { return k_; }
  void resize(int k) { ensure(k); k_ = k; }
  double& operator()(int r, int c) { return a_[(size_t)c * ld_ + r]; }
  double operator()(int r, int c) const { return a_[(size_t)c * ld_ + r]; }
  double* col(int c) { return &a_[(size_t)c * ld_]; }
  const double* col(int c) const { return &a_[(size_t)c * ld_]; }

  void copy_upper_from(const UpperMat& o)
  {
    ensure(o.k_);
    k_ = o.k_;
    for (int c = 0; c < k_; ++c) std::memcpy(col(c), o.col(c), sizeof(double) * (c + 1));
  }
  // SymmMatrix::remove_colrow (symmmatrix.cpp:128-142): drop row and column `rem` of a symmetric matrix
  void remove_colrow(int rem)
  {
    for (int c = rem; c < k_ - 1; ++c)
      for (int r = 0; r <= c; ++r) (*this)(r, c) = r < rem ? (*this)(r, c + 1) : (*this)(r + 1, c + 1);
    --k_;
  }
  // in-place Cholesky A = U'U on the upper triangle; false if not positive definite (dpotrf 'U')
  bool cholesky()
  {
    for (int j = 0; j < k_; ++j) {
      double* cj = col(j);
      double d = cj[j];
      for (int i = 0; i < j; ++i) d -= cj[i] * cj[i];
      if (!(d > 0.0)) return false;
      d = std::sqrt(d);
      cj[j] = d;
      for (int c = j + 1; c < k_; ++c) {
        double* cc = col(c);
        double t = cc[j];
        for (int i = 0; i < j; ++i) t -= cj[i] * cc[i];
        cc[j] = t / d;
      }
    }
    return true;
  }
  // x := U'^-1 x   (forward substitution with the transposed upper factor)
  void solve_transposed(double* x, int k) const
  {
    for (int j = 0; j < k; ++j) {
      const double* cj = col(j);
      double t = x[j];
      for (int i = 0; i < j; ++i) t -= cj[i] * x[i];
      x[j] = t / cj[j];
    }
  }
  // x := U^-1 x   (back substitution)
  void solve(double* x, int k) const
  {
    for (int j = k - 1; j >= 0; --j) {
      const double* cj = col(j);
      x[j] /= cj[j];
      const double t = x[j];
      for (int i = 0; i < j; ++i) x[i] -= t * cj[i];
    }
  }
  // append a column: newcol holds A(0..k-1, k) and A(k,k); tau is added to the diagonal
  // (SymmMatrix::cholesky_update, symmmatrix.cpp:144-164).  false: not positive definite (size still grows).
  bool append(const double* newcol, double tau)
  {
    const int k = k_;
    ensure(k + 1);
    double* c = col(k);
    for (int i = 0; i < k; ++i) c[i] = newcol[i];
    solve_transposed(c, k);
    double ss = 0.0;
    for (int i = 0; i < k; ++i) ss += c[i] * c[i];
    const double d = newcol[k] + tau - ss;
    k_ = k + 1;
    if (d <= 0) return false;
    c[k] = std::sqrt(d);
    return true;
  }
  // delete column `rem` of the factor (SymmMatrix::cholesky_downdate, symmmatrix.cpp:166-195)
  void remove(int rem)
  {
    const int k = k_;
    for (int j = rem; j < k - 1; ++j) {
      double* dst = col(j);
      const double* src = col(j + 1);
      for (int i = 0; i <= j + 1; ++i) dst[i] = src[i];
    }
    for (int j = rem; j < k - 1; ++j) {
      double c, s;
      double a = (*this)(j, j), b = (*this)(j + 1, j);
      rotg(a, b, c, s);
      (*this)(j, j) = a;
      (*this)(j + 1, j) = 0.0;
      for (int q = j + 1; q < k - 1; ++q) {
        const double u = (*this)(j, q), w = (*this)(j + 1, q);
        (*this)(j, q) = c * u + s * w;
        (*this)(j + 1, q) = c * w - s * u;
      }
    }
    k_ = k - 1;
    for (int i = rem; i < k_; ++i)   // keep the diagonal positive (symmmatrix.cpp:184-194)
      if ((*this)(i, i) < 0)
        for (int j = i; j < k_; ++j) (*this)(i, j) = -(*this)(i, j);
  }
  // exchange adjacent columns c, c+1 of the factor and the matching entries of v
  // (SymmMatrix::cholesky_swapadj, symmmatrix.cpp:220-265)
  void swap_adjacent(int c, double* v)
  {
    double* a = col(c);
    double* b = col(c + 1);
    for (int i = 0; i < c + 2; ++i) std::swap(a[i], b[i]);
    b[c + 1] = 0.0;
    double cs, sn;
    rotg(a[c], a[c + 1], cs, sn);
    if (a[c] < 0) { a[c] = -a[c]; sn = -sn; cs = -cs; }
    for (int q = c + 1; q < k_; ++q) {
      double* cq = col(q);
      const double t = cq[c] * sn - cq[c + 1] * cs;
      cq[c] = cq[c + 1] * sn + cq[c] * cs;
      cq[c + 1] = t;
    }
    if (v) {
      const double t = v[c] * sn - v[c + 1] * cs;
      v[c] = v[c + 1] * sn + v[c] * cs;
      v[c + 1] = t;
    }
  }

 private:
  // BLAS drotg (as called from dchex.f:229 and symmmatrix.cpp:237): a := r, b := z
  static void rotg(double& a, double& b, double& c, double& s)
  {
    const double roe = std::fabs(a) > std::fabs(b) ? a : b;
    const double scale = std::fabs(a) + std::fabs(b);
    double r, z;
    if (scale == 0.0) { c = 1.0; s = 0.0; r = 0.0; z = 0.0; }
    else {
      const double ta = a / scale, tb = b / scale;
      r = scale * std::sqrt(ta * ta + tb * tb);
      if (roe < 0) r = -r;
      c = a / r;
      s = b / r;
      z = 1.0;
      if (std::fabs(a) > std::fabs(b)) z = s;
      if (std::fabs(b) >= std::fabs(a) && c != 0.0) z = 1.0 / c;
    }
    a = r;
    b = z;
  }
  void ensure(int k)
  {
    if (k <= ld_) return;
    int nld = ld_;
    while (nld < k) nld *= 2;
    std::vector<double> na((size_t)nld * nld, 0.0);
    for (int c = 0; c < k_; ++c) std::memcpy(&na[(size_t)c * nld], col(c), sizeof(double) * (c + 1));
    a_.swap(na);
    ld_ = nld;
  }
  int ld_, k_;
  std::vector<double> a_;
};

}  // namespace bmg
