/* atom_meso.h -- MesoAtom, constructed by name in src/lammps.cpp:532-568 when LAMMPS runs with -meso on.
   Reference: UM/atom_meso.h:11 (device atom store owner; here the store lives behind the C ABI).
   Host-side behaviour is the stock Atom: read_data, velocity, thermo, dump and restart keep working
   on host arrays, which ModifiedVerlet uploads at setup and refreshes at every output step. */
#ifndef LMP_MESO_ATOM
#define LMP_MESO_ATOM

#include "atom.h"

namespace LAMMPS_NS {

class MesoAtom : public Atom {
 public:
  MesoAtom(class LAMMPS *lmp) : Atom(lmp) {}
  virtual ~MesoAtom() {}
};

}

#endif
