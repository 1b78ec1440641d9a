// feMat<dim>: abstract base of the matrix-free operators (FEM/include/feMat.h:18-116).
#ifndef DKT_HOST_FEMAT_H
#define DKT_HOST_FEMAT_H

#include "oda.h"

template <unsigned int dim>
class feMat
{
protected:
  ot::DA<dim> *m_uiOctDA;  // not owned (feMat.h:45-50)
  double m_uiPtMin[dim], m_uiPtMax[dim];

public:
  feMat(ot::DA<dim> *da) : m_uiOctDA(da)
  {
    for (unsigned d = 0; d < dim; d++) { m_uiPtMin[d] = 0.0; m_uiPtMax[d] = 1.0; }
  }
  virtual ~feMat() {}
  virtual void matVec(const VECType *in, VECType *out, double scale = 1.0) = 0;
  void setProblemDimensions(const double *pt_min, const double *pt_max)
  {
    for (unsigned d = 0; d < dim; d++) { m_uiPtMin[d] = pt_min[d]; m_uiPtMax[d] = pt_max[d]; }
  }
};
#endif
