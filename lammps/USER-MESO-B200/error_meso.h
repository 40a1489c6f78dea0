/* error_meso.h -- MesoError, constructed by name in src/lammps.cpp:532-568 when LAMMPS runs with -meso on.
   Reference: UM/error_meso.h.
   Host-side behaviour is the stock Error: read_data, velocity, thermo, dump and restart keep working
   on host arrays, which ModifiedVerlet uploads at setup and refreshes at every output step. */
#ifndef LMP_MESO_ERROR
#define LMP_MESO_ERROR

#include "error.h"

namespace LAMMPS_NS {

class MesoError : public Error {
 public:
  MesoError(class LAMMPS *lmp) : Error(lmp) {}
  virtual ~MesoError() {}
};

}

#endif
