/* fix_addforce_meso.h -- fix ID group addforce/meso fx fy fz   (UM/fix_addforce_meso.h, UM/fix_addforce_meso.cu:20-108) */
#ifdef FIX_CLASS

FixStyle(addforce/meso,MesoFixAddForce)

#else

#ifndef LMP_MESO_FIX_ADD_FORCE
#define LMP_MESO_FIX_ADD_FORCE

#include "fix_resident_meso.h"

namespace LAMMPS_NS {

class MesoFixAddForce : public MesoFixResident {
 public:
  MesoFixAddForce(class LAMMPS *, int, char **);
  virtual int setmask();
 protected:
  double fx, fy, fz;
  virtual int register_fix(meso_ctx *);
};

}

#endif
#endif
