/* atom_vec_dpd_bond_meso.h -- atom_style dpd/bond/meso (UM/atom_vec_dpd_bond_meso.h:3,15-47).
   Per-atom fields of stock `bond` (molecule, num_bond, bond_type, bond_atom, special lists) next to x, v, f;
   velocities travel with ghosts.  The device copy of the bond table ({partner tag, type} per slot, column-major)
   is owned by the library and rides its reorder and migration; MesoDevice re-aligns the host-only arrays
   (molecule, bond and special tables) with the device order by tag whenever atoms come back. */
#ifdef ATOM_CLASS

AtomStyle(dpd/bond/meso,AtomVecDPDBond)

#else

#ifndef LMP_MESO_ATOM_VEC_DPD_BOND
#define LMP_MESO_ATOM_VEC_DPD_BOND

#include "atom_vec_bond.h"

namespace LAMMPS_NS {

class AtomVecDPDBond : public AtomVecBond {
 public:
  AtomVecDPDBond(class LAMMPS *lmp) : AtomVecBond(lmp)
  {
    cudable = 1;
    comm_x_only = 0;
  }
  virtual ~AtomVecDPDBond() {}
};

}

#endif
#endif
