#include "string.h"
#include "fix_styles_meso.h"
#include "error.h"

using namespace LAMMPS_NS;
using namespace FixConst;

/* argument grammar and messages of UM/fix_solid_bound_meso.cu:24-46 */
MesoFixSolidBound::MesoFixSolidBound(LAMMPS *lmp, int narg, char **arg) : MesoFixResident(lmp,narg,arg)
{
  if (narg < 4) error->all(FLERR,"Illegal fix MesoFixSolidBound command");
  x = y = z = false;
  force_kernel = 0;
  for (int i = 0; i < narg; i++) {
    if (!strcmp(arg[i],"x")) x = true;
    else if (!strcmp(arg[i],"y")) y = true;
    else if (!strcmp(arg[i],"z")) z = true;
    else if (!strcmp(arg[i],"rho5rc1s1")) force_kernel = 1;
  }
  if (!x && !y && !z) error->all(FLERR,"Incomplete fix wall command: dimension unspecified");
  if (force_kernel == 0) error->all(FLERR,"Incomplete fix wall command: force kernel unspecified");
  nevery = 1;
}

int MesoFixSolidBound::setmask() { return POST_FORCE | PRE_EXCHANGE | END_OF_STEP; }

int MesoFixSolidBound::register_fix(meso_ctx *ctx)
{
  return meso_fix_solid_bound(ctx,groupbit,(x ? 1 : 0) | (y ? 2 : 0) | (z ? 4 : 0),force_kernel);
}
