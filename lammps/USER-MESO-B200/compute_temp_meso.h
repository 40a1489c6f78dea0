/* compute_temp_meso.h -- compute ID group temp/meso   (UM/compute_temp_meso.h, .cu:18-101)
   T = sum_i m_i v_i^2 * mvv2e / (dof * k_B) over the group, reduced on the device (meso_compute_ke). */
#ifdef COMPUTE_CLASS

ComputeStyle(temp/meso,MesoComputeTemp)

#else

#ifndef LMP_MESO_COMPUTE_TEMP
#define LMP_MESO_COMPUTE_TEMP

#include "compute.h"
#include "meso_bridge.h"

namespace LAMMPS_NS {

class MesoComputeTemp : public Compute, protected MesoBridge {
 public:
  MesoComputeTemp(class LAMMPS *, int, char **);
  virtual void init() {}
  virtual void setup();
  virtual double compute_scalar();

 protected:
  int fix_dof;
  double tfactor;
  virtual void dof_compute();
};

}

#endif
#endif
