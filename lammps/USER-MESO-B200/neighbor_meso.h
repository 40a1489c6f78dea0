/* neighbor_meso.h -- MesoNeighbor, constructed by src/lammps.cpp:543 with -meso on.
   Reference: UM/neighbor_meso.h:27 (device binning + list build behind Neighbor::build).
   Cell binning, the 2-level reorder and the tile-transposed neighbor table are built inside
   the library (meso_rebuild); the host Neighbor keeps its cutoff bookkeeping (cutneighmax,
   skin, every) because Comm::setup and the pair styles read it, and builds no host lists. */
#ifndef LMP_MESO_NEIGHBOR
#define LMP_MESO_NEIGHBOR

#include "neighbor.h"

namespace LAMMPS_NS {

class MesoNeighbor : public Neighbor {
 public:
  MesoNeighbor(class LAMMPS *lmp) : Neighbor(lmp) {}
  virtual ~MesoNeighbor() {}
  // re-neighboring is decided by the library on the deck's fixed cadence (meso_neighbor_decide)
  virtual int check_distance() { return 0; }
};

}

#endif
