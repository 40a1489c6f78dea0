/* fix_wall_meso.h -- fix ID group wall/meso [x] [y] [z] d <d> f <f>   (UM/fix_wall_meso.h, UM/fix_wall_meso.cu:20-238)
   Soft erfc wall force within d of the box faces (post_force) and bounce-forward at the faces
   (pre_exchange, end_of_step). */
#ifdef FIX_CLASS

FixStyle(wall/meso,MesoFixWall)

#else

#ifndef LMP_MESO_FIX_WALL
#define LMP_MESO_FIX_WALL

#include "fix_resident_meso.h"

namespace LAMMPS_NS {

class MesoFixWall : public MesoFixResident {
 public:
  MesoFixWall(class LAMMPS *, int, char **);
  virtual int setmask();
  virtual void pre_exchange() { bounce(); }
  virtual void end_of_step() { bounce(); }
 protected:
  bool x, y, z;
  double d, f;
  virtual int register_fix(meso_ctx *);
};

}

#endif
#endif
