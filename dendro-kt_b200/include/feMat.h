// feMat<dim>: abstract base of the matrix-free operators (FEM/include/feMat.h:18-116).
#ifndef DKT_HOST_FEMAT_H
#define DKT_HOST_FEMAT_H

#include "oda.h"
#include "point.h"

template <unsigned int dim>
class feMat
{
protected:
  static constexpr unsigned int m_uiDim = dim;
  ot::DA<dim> *m_uiOctDA;            // not owned (feMat.h:45-50)
  Point<dim> m_uiPtMin, m_uiPtMax;   // problem domain (feMat.h:33-37); the unit cube unless set

public:
  feMat(ot::DA<dim> *da) : m_uiOctDA(da), m_uiPtMin(0.0), m_uiPtMax(1.0) {}
  virtual ~feMat() {}
  virtual void matVec(const VECType *in, VECType *out, double scale = 1.0) = 0;
  inline void setProblemDimensions(const Point<dim> &pt_min, const Point<dim> &pt_max)
  {
    m_uiPtMin = pt_min;
    m_uiPtMax = pt_max;
  }
  void setProblemDimensions(const double *pt_min, const double *pt_max)
  {
    m_uiPtMin = Point<dim>(pt_min);
    m_uiPtMax = Point<dim>(pt_max);
  }
};
#endif
