/* timer_meso.h -- MesoTimer, constructed by name in src/lammps.cpp:532-568 when LAMMPS runs with -meso on.
   Reference: UM/timer_meso.h.
   Host-side behaviour is the stock Timer: read_data, velocity, thermo, dump and restart keep working
   on host arrays, which ModifiedVerlet uploads at setup and refreshes at every output step. */
#ifndef LMP_MESO_TIMER
#define LMP_MESO_TIMER

#include "timer.h"

namespace LAMMPS_NS {

class MesoTimer : public Timer {
 public:
  MesoTimer(class LAMMPS *lmp) : Timer(lmp) {}
  virtual ~MesoTimer() {}
};

}

#endif
