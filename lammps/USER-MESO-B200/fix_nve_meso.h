/* fix_nve_meso.h -- fix ID group nve/meso   (UM/fix_nve_meso.h, UM/fix_nve_meso.cu:29-205)
   Velocity-Verlet halves on the device SoA store: dtf = 0.5*dt*ftm2v, dtv = dt, group-mask test.
   ModifiedVerlet fuses both halves into neighbouring kernels when this is the only integrating fix. */
#ifdef FIX_CLASS

FixStyle(nve/meso,FixNVEMeso)

#else

#ifndef LMP_MESO_FIX_NVE
#define LMP_MESO_FIX_NVE

#include "fix.h"
#include "meso_bridge.h"

namespace LAMMPS_NS {

class FixNVEMeso : public Fix, protected MesoBridge {
 public:
  FixNVEMeso(class LAMMPS *, int, char **);
  virtual int setmask();
  virtual void init();
  virtual void initial_integrate(int);
  virtual void final_integrate();
  virtual void reset_dt();
  int group_bit() const { return groupbit; }
};

}

#endif
#endif
