/* comm_meso.h -- MesoComm, constructed by name in src/lammps.cpp:532-568 when LAMMPS runs with -meso on.
   Reference: UM/comm_meso.h:11 (host MPI halo; here halo and migration run on the device, meso_rebuild / meso_forward_comm).
   Host-side behaviour is the stock Comm: read_data, velocity, thermo, dump and restart keep working
   on host arrays, which ModifiedVerlet uploads at setup and refreshes at every output step. */
#ifndef LMP_MESO_COMM
#define LMP_MESO_COMM

#include "comm.h"

namespace LAMMPS_NS {

class MesoComm : public Comm {
 public:
  MesoComm(class LAMMPS *lmp) : Comm(lmp) {}
  virtual ~MesoComm() {}
};

}

#endif
