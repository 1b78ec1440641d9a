// feVec<dim>: abstract base of the right-hand-side assemblers (FEM/include/feVec.h:17-97).
#ifndef DKT_HOST_FEVEC_H
#define DKT_HOST_FEVEC_H

#include "oda.h"
#include "point.h"

template <unsigned int dim>
class feVec
{
protected:
  static constexpr unsigned int m_uiDim = dim;
  ot::DA<dim> *m_uiOctDA;            // not owned
  Point<dim> m_uiPtMin, m_uiPtMax;   // problem domain; the unit cube unless set

public:
  feVec(ot::DA<dim> *da) : m_uiOctDA(da), m_uiPtMin(0.0), m_uiPtMax(1.0) {}
  virtual ~feVec() {}
  virtual void computeVec(const VECType *in, VECType *out, double scale = 1.0) = 0;
  inline void setProblemDimensions(const Point<dim> &pt_min, const Point<dim> &pt_max)
  {
    m_uiPtMin = pt_min;
    m_uiPtMax = pt_max;
  }
};
#endif
