/* fix_poiseuille_meso.h -- fix ID group pois/meso <dim_ortho> <dim_force> <strength> [bisect_frac]
   (UM/fix_poiseuille_meso.h, UM/fix_poiseuille_meso.cu:20-113): body force +strength on one side of the bisection
   plane across dim_ortho and -strength on the other (periodic reverse-Poiseuille driving). */
#ifdef FIX_CLASS

FixStyle(pois/meso,MesoFixPoiseuille)

#else

#ifndef LMP_MESO_FIX_POISEUILLE
#define LMP_MESO_FIX_POISEUILLE

#include "fix_resident_meso.h"

namespace LAMMPS_NS {

class MesoFixPoiseuille : public MesoFixResident {
 public:
  MesoFixPoiseuille(class LAMMPS *, int, char **);
  virtual int setmask();
 protected:
  int dim_ortho, dim_force;
  double strength, bisect_frac;
  virtual int register_fix(meso_ctx *);
};

}

#endif
#endif
