/* fix_solid_bound_meso.h -- fix ID group solid_bound/meso [x] [y] [z] rho5rc1s1
   (UM/fix_solid_bound_meso.h, UM/fix_solid_bound_meso.cu:20-234): polynomial wall force fitted for
   rho = 5, rc = 1, s = 1 within one cutoff of the box faces + bounce-forward. */
#ifdef FIX_CLASS

FixStyle(solid_bound/meso,MesoFixSolidBound)

#else

#ifndef LMP_MESO_FIX_SOLID_BOUND
#define LMP_MESO_FIX_SOLID_BOUND

#include "fix_resident_meso.h"

namespace LAMMPS_NS {

class MesoFixSolidBound : public MesoFixResident {
 public:
  MesoFixSolidBound(class LAMMPS *, int, char **);
  virtual int setmask();
  virtual void pre_exchange() { bounce(); }
  virtual void end_of_step() { bounce(); }
 protected:
  bool x, y, z;
  int force_kernel;              // 0 unspecified, 1 rho5rc1s1
  virtual int register_fix(meso_ctx *);
};

}

#endif
#endif
