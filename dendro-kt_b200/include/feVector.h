// feVector<LeafT,dim>::computeVec over libdkt.so (FEM/include/feVector.h:17-168 of the reference): the
// right-hand-side assembly b = M f uses the same traversal as the matvec with the callback
// elementalComputeVec(in, out, coords, scale); it is probed and verified exactly like feMatrix.
#ifndef DKT_HOST_FEVECTOR_H
#define DKT_HOST_FEVECTOR_H

#include "feMatrix.h"
#include "feVec.h"

template <typename LeafT, unsigned int dim>
class feVector : public feVec<dim>
{
  // adaptor: expose elementalComputeVec as an elementalMatVec
  struct Adaptor : public feMatrix<Adaptor, dim>
  {
    feVector *owner;
    Adaptor(ot::DA<dim> *da, feVector *o) : feMatrix<Adaptor, dim>(da, 1), owner(o) {}
    virtual void elementalMatVec(const VECType *in, VECType *out, double *coords, double scale)
    {
      owner->elementalComputeVec(in, out, coords, scale);
    }
    bool preMatVec(const VECType *in, VECType *out, double scale) { return owner->asLeaf().preComputeVec(in, out, scale); }
    bool postMatVec(const VECType *in, VECType *out, double scale) { return owner->asLeaf().postComputeVec(in, out, scale); }
  };
  Adaptor m_impl;

protected:
  using feVec<dim>::m_uiOctDA;
  unsigned int m_uiDof;

public:
  feVector(ot::DA<dim> *da, unsigned int dof = 1) : feVec<dim>(da), m_impl(da, this), m_uiDof(dof) {}
  virtual ~feVector() {}
  virtual void elementalComputeVec(const VECType *in, VECType *out, double *coords, double scale) = 0;
  LeafT &asLeaf() { return static_cast<LeafT &>(*this); }
  bool preComputeVec(const VECType *, VECType *, double = 1.0) { return false; }
  bool postComputeVec(const VECType *, VECType *, double = 1.0) { return false; }
  virtual void computeVec(const VECType *in, VECType *out, double scale = 1.0) { m_impl.matVec(in, out, scale); }
};
#endif
