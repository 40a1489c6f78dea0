/* compute_styles_meso.h -- the two compute styles of the package.

     temp/meso   MesoComputeTemp   UM/compute_temp_meso.{h,cu}:18-101
                 T = sum_i m_i v_i^2 * mvv2e / (dof * k_B) over the group, reduced on the device (meso_compute_ke)
     pe/meso     MesoComputePE     UM/compute_pe_meso.{h,cu}:66-125
                 src/output.cpp:62-69 creates it as thermo_pe whenever the package is compiled in, also with -meso off.
                 Pair (dpd: 1/2 a0 cut wc^2 per pair) + bond energy of the device-resident styles, reduced on the device when
                 the force evaluation of this step tallied energy; keywords pair, bond, thermo. */
#ifdef COMPUTE_CLASS

ComputeStyle(temp/meso,MesoComputeTemp)
ComputeStyle(pe/meso,MesoComputePE)

#else
#ifndef LMP_MESO_B200_COMPUTE_STYLES_H
#define LMP_MESO_B200_COMPUTE_STYLES_H

#include "compute.h"
#include "meso_bridge.h"

namespace LAMMPS_NS {

class MesoComputeTemp : public Compute, protected MesoBridge {
 public:
  MesoComputeTemp(class LAMMPS *, int, char **);
  void init() {}
  void setup();
  double compute_scalar();
 protected:
  int fix_dof;
  double tfactor;
  void dof_compute();
};

class MesoComputePE : public Compute {
 public:
  MesoComputePE(class LAMMPS *, int, char **);
  void init() {}
  double compute_scalar();
 private:
  int pairflag, bondflag, thermoflag;
};

}

#endif
#endif
