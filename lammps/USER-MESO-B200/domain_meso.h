/* domain_meso.h -- MesoDomain, constructed by name in src/lammps.cpp:532-568 when LAMMPS runs with -meso on.
   Reference: UM/domain_meso.h:10 (OpenMP pbc; here the wrap is fused into the device reorder key kernel).
   Host-side behaviour is the stock Domain: read_data, velocity, thermo, dump and restart keep working
   on host arrays, which ModifiedVerlet uploads at setup and refreshes at every output step. */
#ifndef LMP_MESO_DOMAIN
#define LMP_MESO_DOMAIN

#include "domain.h"

namespace LAMMPS_NS {

class MesoDomain : public Domain {
 public:
  MesoDomain(class LAMMPS *lmp) : Domain(lmp) {}
  virtual ~MesoDomain() {}
};

}

#endif
