/* compute_pe_meso.h -- compute ID group pe/meso   (UM/compute_pe_meso.h, .cu:66-125)
   src/output.cpp:62-69 creates it as thermo_pe whenever the package is compiled in, also with -meso off.
   Potential energy of the device-resident pair style (dpd: 1/2 a0 cut wc^2 per pair), reduced on the device
   when the force evaluation of this step tallied energy. */
#ifdef COMPUTE_CLASS

ComputeStyle(pe/meso,MesoComputePE)

#else

#ifndef LMP_MESO_COMPUTE_PE
#define LMP_MESO_COMPUTE_PE

#include "compute.h"

namespace LAMMPS_NS {

class MesoComputePE : public Compute {
 public:
  MesoComputePE(class LAMMPS *, int, char **);
  virtual void init() {}
  virtual double compute_scalar();

 private:
  int pairflag, bondflag, thermoflag;
};

}

#endif
#endif
