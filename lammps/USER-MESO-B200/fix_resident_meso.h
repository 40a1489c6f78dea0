/* fix_resident_meso.h -- base of the fixes whose whole effect lives in the library's device-resident fix list
   (wall/meso, solid_bound/meso, addforce/meso, pois/meso).  A fix registers itself with the context in init()
   (ModifiedVerlet::init() has cleared the list just before: LAMMPS::init runs Update::init ahead of Modify::init)
   and keeps the handle.  With only such fixes next to one nve/meso the fused run loop (meso_run) applies them
   itself; otherwise the phase hooks below are called once per step by Modify. */
#ifndef LMP_MESO_FIX_RESIDENT_H
#define LMP_MESO_FIX_RESIDENT_H

#include "fix.h"
#include "meso_bridge.h"

namespace LAMMPS_NS {

class MesoFixResident : public Fix, protected MesoBridge {
 public:
  MesoFixResident(class LAMMPS *lmp, int narg, char **arg) : Fix(lmp,narg,arg), MesoBridge(lmp), handle(-1) {}
  virtual void init();
  virtual void setup(int) {}            // Fix::setup -> post_force of the reference: meso_setup has applied the whole list
  virtual void post_force(int);
 protected:
  int handle;
  const char *who;
  virtual int register_fix(meso_ctx *) = 0;
  void bounce();
};

}

#endif
