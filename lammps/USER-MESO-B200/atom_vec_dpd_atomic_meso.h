/* atom_vec_dpd_atomic_meso.h -- atom_style dpd/atomic/meso (UM/atom_vec_dpd_atomic_meso.h:3,15-33).
   Per-atom fields: tag, type, mask, image, x, v, f -- exactly stock `atomic`; velocities travel
   with ghosts (comm_x_only = 0) because the DPD drag and the pair RNG signature need them.
   The device SoA copy of these fields is owned by the library (A1 in SURVEY.md s8). */
#ifdef ATOM_CLASS

AtomStyle(dpd/atomic/meso,AtomVecDPDAtomic)

#else

#ifndef LMP_MESO_ATOM_VEC_DPD_ATOMIC
#define LMP_MESO_ATOM_VEC_DPD_ATOMIC

#include "atom_vec_atomic.h"

namespace LAMMPS_NS {

class AtomVecDPDAtomic : public AtomVecAtomic {
 public:
  AtomVecDPDAtomic(class LAMMPS *lmp) : AtomVecAtomic(lmp)
  {
    cudable = 1;
    comm_x_only = 0;
  }
  virtual ~AtomVecDPDAtomic() {}
};

}

#endif
#endif
