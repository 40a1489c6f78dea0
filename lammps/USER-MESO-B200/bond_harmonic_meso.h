/* bond_harmonic_meso.h -- bond_style harmonic/meso (UM/bond_harmonic_meso.h, UM/bond_harmonic_meso.cu:34-170).
   Coefficient grammar, restart records and single() are stock BondHarmonic's (E = K (r - r0)^2); compute() runs on the
   device over the library's per-atom bond table (newton_bond off: every atom sums its own entries). */
#ifdef BOND_CLASS

BondStyle(harmonic/meso,MesoBondHarmonic)

#else

#ifndef LMP_MESO_BOND_HARMONIC
#define LMP_MESO_BOND_HARMONIC

#include "bond_harmonic.h"
#include "meso_bridge.h"

namespace LAMMPS_NS {

class MesoBondHarmonic : public BondHarmonic, protected MesoBridge {
 public:
  MesoBondHarmonic(class LAMMPS *lmp) : BondHarmonic(lmp), MesoBridge(lmp) {}
  virtual void init_style();
  virtual void compute(int, int);
  void push_coeff();                       // bond_coeff table -> device (MesoBondHarmonic::alloc_coeff)
  void tally_from_device(int, int);        // energy of the last flagged evaluation -> Bond::energy
};

}

#endif
#endif
