/* meso_bridge.h -- what every style of this package shares: access to the C-ABI context
   that MesoDevice owns, and conversion of a library status code into a LAMMPS error.
   Plays the part of MesoPointers in the reference (UM/pointers_meso.h:30-61), but hands
   out nothing except the opaque meso_ctx: the styles own no device memory. */
#ifndef LMP_MESO_BRIDGE_H
#define LMP_MESO_BRIDGE_H

#include "meso_b200.h"

namespace LAMMPS_NS {

class LAMMPS;
class MesoDevice;

class MesoBridge {
 public:
  MesoBridge(LAMMPS *l) : mlmp(l) {}

 protected:
  LAMMPS *mlmp;
  MesoDevice *mdev(const char *who) const;   // error->all if LAMMPS runs with -meso off
  meso_ctx *mctx(const char *who) const;
  void mcheck(int rc, const char *file, int line) const;
};

}

#define MESO_CALL(expr) mcheck((expr), __FILE__, __LINE__)

#endif
