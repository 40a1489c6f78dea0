/* run_style_meso.h -- run_style mvv/meso | verlet/meso   (UM/mvv_meso.h:3-4, UM/mvv_meso.cu:79-435)
   Drives one DPD time step in the order of ModifiedVerlet::run:
     initial_integrate -> [rebuild every N steps: wrap, migrate, reorder, ghosts, neighbor table | halo refresh]
     -> force clear -> pair (bulk overlapped with the halo, then border) -> final_integrate -> output.
   All phases execute inside the library; consecutive steps that need no energy/virial tally and whose only
   integrating fix is nve/meso are handed over in one meso_run() call (fused kernels, no host work per step). */
#ifdef INTEGRATE_CLASS

IntegrateStyle(mvv/meso,ModifiedVerlet)
IntegrateStyle(verlet/meso,ModifiedVerlet)

#else

#ifndef LMP_MESO_B200_RUN_STYLE_H
#define LMP_MESO_B200_RUN_STYLE_H

#include "integrate.h"
#include "meso_bridge.h"

namespace LAMMPS_NS {

class ModifiedVerlet : public Integrate, protected MesoBridge {
 public:
  ModifiedVerlet(class LAMMPS *, int, char **);
  // Integrate interface (src/integrate.h): all of it runs on device-resident atoms
  void init(); void setup(); void setup_minimal(int);
  void run(int); void cleanup(); void reset_dt();

 protected:
  int fused_groupbit;         // group of the single nve/meso fix, or -1 when the fix list cannot be fused
  class MesoPairDPD *dpd;

  void device_setup(int outflag);
  void step_by_phases(bigint ntimestep);
  void flush(int &pending);
  bool thermo_reads_device_only();
  void force_clear();
};

}

#endif
#endif
