// TEST INFRASTRUCTURE: the evaluation semantics oracle/eigen_shim promises (see Eigen/Core), on inputs where the
// order of the floating-point operations is visible in the result.  Built and run by tests/test_eigen_shim.py.
#include <Eigen/Core>
#include <Eigen/Sparse>
#include <cstdio>
#include <vector>

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); ++fails; } } while (0)

int main() {
  using namespace Eigen;
  // (a - b) + c, coefficient by coefficient, left to right: 1e16 - 1e16 + 1 = 1, whereas 1e16 + (-1e16 + 1) = 0
  {
    MatrixXd a(1, 1), b(1, 1), c(1, 1);
    a(0, 0) = 1e16; b(0, 0) = 1e16; c(0, 0) = 1.0;
    MatrixXd r = a - b + c;
    CHECK(r(0, 0) == 1.0);
  }
  // dense * sparse: each coefficient is summed over the rows of the sparse column in ASCENDING order from zero,
  // whatever the order of the triplets: (0 + 1e16) + 1 - 1e16 -> 0 ; another order would give 1 or 2
  {
    SparseMatrix<double> m(3, 1);
    std::vector<Triplet<double>> t = {{2, 0, -1.0}, {0, 0, 1.0}, {1, 0, 1.0}};   // rows 2, 0, 1: sorted on construction
    m.setFromTriplets(t.begin(), t.end());
    m.makeCompressed();
    MatrixXd c(1, 3);
    c(0, 0) = 1e16; c(0, 1) = 1.0; c(0, 2) = 1e16;
    MatrixXd r = c * m;   // ((0 + 1e16*1) + 1*1) + 1e16*(-1) = (1e16 + 1) - 1e16 = 0 (1e16 + 1 rounds to 1e16)
    CHECK(r.rows() == 1 && r.cols() == 1 && r(0, 0) == 0.0);
    CHECK(m.nonZeros() == 3 && m.innerIndexPtr()[0] == 0 && m.innerIndexPtr()[1] == 1 && m.innerIndexPtr()[2] == 2);
  }
  // duplicate triplets are summed (in input order) into one entry
  {
    SparseMatrix<double> m(2, 2);
    std::vector<Triplet<double>> t = {{1, 0, 0.25}, {0, 1, 2.0}, {1, 0, 0.5}};
    m.setFromTriplets(t.begin(), t.end());
    CHECK(m.nonZeros() == 2 && m.valuePtr()[0] == 0.75 && m.valuePtr()[1] == 2.0);
  }
  // dense * diagonal scales column j by d(j); Map writes through to the mapped storage; += accumulates in place
  {
    double buf[4] = {1, 2, 3, 4};   // column-major 2 x 2: [1 3; 2 4]
    Map<MatrixXd> c(buf, 2, 2);
    DiagonalMatrix<double, Dynamic> d(2);
    d.diagonal()(0) = 10.0; d.diagonal()(1) = 100.0;
    MatrixXd r = c * d;
    CHECK(r(0, 0) == 10.0 && r(1, 0) == 20.0 && r(0, 1) == 300.0 && r(1, 1) == 400.0);
    MatrixXd acc(2, 2); acc.setZero();
    acc.noalias() += 0.5 * r;
    c.noalias() = acc * d;
    CHECK(buf[0] == 50.0 && buf[3] == 20000.0);
  }
  // Array semantics: colwise() * vector, coefficient-wise products and quotients, pow, row assignment from a transpose
  {
    ArrayXXd g(2, 3), l(2, 3), kla(2, 3);
    for (int j = 0; j < 3; ++j) { g(0, j) = 1 + j; g(1, j) = 10 * (1 + j); l(0, j) = 0.5; l(1, j) = 1.0; kla(0, j) = 0.0; kla(1, j) = 2.0; }
    ArrayXd h(2); h(0) = 0.0; h(1) = 0.1;
    DiagonalMatrix<double, Dynamic> v(3);
    for (int j = 0; j < 3; ++j) v.diagonal()(j) = 2.0 * (j + 1);
    MatrixXd mtr = (kla * (g.colwise() * h - l)).matrix() * MatrixXd(v);
    CHECK(mtr(0, 1) == 0.0 && mtr(1, 0) == (2.0 * (10 * 0.1 - 1.0)) * 2.0 && mtr(1, 2) == (2.0 * (30 * 0.1 - 1.0)) * 6.0);
    ArrayXd e(3); e(0) = 16.0; e(1) = 81.0; e(2) = 1.0;
    kla.row(1) = (0.5 * e.pow(0.25) * 2.0).transpose();
    CHECK(kla(1, 0) == 2.0 && kla(1, 1) == 3.0 && kla(1, 2) == 1.0 && kla(0, 0) == 0.0);
    auto q = e / (e + e);
    CHECK(q(0, 0) == 0.5 && q.rows() == 3);
  }
  std::printf(fails ? "eigen_shim selftest: %d FAILED\n" : "eigen_shim selftest ok\n", fails);
  return fails ? 1 : 0;
}
