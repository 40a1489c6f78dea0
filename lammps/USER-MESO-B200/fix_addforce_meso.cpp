#include "stdlib.h"
#include "fix_styles_meso.h"
#include "error.h"

using namespace LAMMPS_NS;
using namespace FixConst;

MesoFixAddForce::MesoFixAddForce(LAMMPS *lmp, int narg, char **arg) : MesoFixResident(lmp,narg,arg)
{
  if (narg < 6) error->all(FLERR,"Illegal fix addforce/meso command");
  fx = atof(arg[3]);
  fy = atof(arg[4]);
  fz = atof(arg[5]);
}

int MesoFixAddForce::setmask() { return POST_FORCE; }

int MesoFixAddForce::register_fix(meso_ctx *ctx) { return meso_fix_addforce(ctx,groupbit,fx,fy,fz); }
