#include "stdlib.h"
#include "string.h"
#include "fix_styles_meso.h"
#include "error.h"

using namespace LAMMPS_NS;
using namespace FixConst;

/* argument grammar and messages of UM/fix_wall_meso.cu:24-46 */
MesoFixWall::MesoFixWall(LAMMPS *lmp, int narg, char **arg) : MesoFixResident(lmp,narg,arg)
{
  if (narg < 6) error->all(FLERR,"Illegal fix MesoFixWall command");
  d = 0.0;
  f = 0.0;
  x = y = z = false;
  for (int i = 0; i < narg; i++) {
    if (!strcmp(arg[i],"d")) {
      if (++i >= narg) error->all(FLERR,"Incomplete fix wall command after 'd'");
      d = atof(arg[i]);
    } else if (!strcmp(arg[i],"f")) {
      if (++i >= narg) error->all(FLERR,"Incomplete fix wall command after 'f'");
      f = atof(arg[i]);
    } else if (!strcmp(arg[i],"x")) x = true;
    else if (!strcmp(arg[i],"y")) y = true;
    else if (!strcmp(arg[i],"z")) z = true;
  }
  if (!x && !y && !z) error->all(FLERR,"Incomplete fix wall command: insufficient arguments");
  nevery = 1;
}

int MesoFixWall::setmask() { return POST_FORCE | PRE_EXCHANGE | END_OF_STEP; }

int MesoFixWall::register_fix(meso_ctx *ctx)
{
  return meso_fix_wall(ctx,groupbit,(x ? 1 : 0) | (y ? 2 : 0) | (z ? 4 : 0),d,f);
}
