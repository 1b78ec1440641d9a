// Point<dim>: the small coordinate tuple of the reference's include/point.h (x(), y(), z(), arithmetic), as far as
// the FEM layer uses it (FEM/include/feMat.h:33-37, FEM/examples/src/heatMat.cpp:58-66).
#ifndef DKT_HOST_POINT_H
#define DKT_HOST_POINT_H

#include <array>
#include <cmath>

template <unsigned int dim>
class Point
{
  std::array<double, dim> m_c{};

public:
  Point() {}
  explicit Point(const double *c) { for (unsigned d = 0; d < dim; d++) m_c[d] = c[d]; }
  explicit Point(const std::array<double, dim> &c) : m_c(c) {}
  explicit Point(double s) { m_c.fill(s); }
  Point(double x, double y, double z = 0.0)
  {
    const double v[3] = {x, y, z};
    for (unsigned d = 0; d < dim && d < 3; d++) m_c[d] = v[d];
  }
  double x(unsigned d) const { return m_c[d]; }
  double x() const { return m_c[0]; }
  double y() const { return dim > 1 ? m_c[dim > 1 ? 1 : 0] : 0.0; }
  double z() const { return dim > 2 ? m_c[dim > 2 ? 2 : 0] : 0.0; }
  double &operator[](unsigned d) { return m_c[d]; }
  double operator[](unsigned d) const { return m_c[d]; }
  Point operator+(const Point &o) const { Point r; for (unsigned d = 0; d < dim; d++) r.m_c[d] = m_c[d] + o.m_c[d]; return r; }
  Point operator-(const Point &o) const { Point r; for (unsigned d = 0; d < dim; d++) r.m_c[d] = m_c[d] - o.m_c[d]; return r; }
  Point operator*(double s) const { Point r; for (unsigned d = 0; d < dim; d++) r.m_c[d] = m_c[d] * s; return r; }
  bool operator==(const Point &o) const { return m_c == o.m_c; }
  double abs() const { double a = 0.0; for (unsigned d = 0; d < dim; d++) a += m_c[d] * m_c[d]; return std::sqrt(a); }
};
#endif
