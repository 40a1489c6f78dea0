#include "ctype.h"
#include "stdlib.h"
#include "fix_styles_meso.h"
#include "error.h"

using namespace LAMMPS_NS;
using namespace FixConst;

static int parse_dim(const char *s)
{
  if (isdigit(s[0])) return atoi(s);
  return s[0] == 'x' ? 0 : (s[0] == 'y' ? 1 : (s[0] == 'z' ? 2 : 0));   // unknown letters map to 0 like std::map's default
}

MesoFixPoiseuille::MesoFixPoiseuille(LAMMPS *lmp, int narg, char **arg) : MesoFixResident(lmp,narg,arg)
{
  if (narg < 6) error->all(FLERR,"Illegal fix CUDAPoiseuille command");
  dim_ortho = parse_dim(arg[3]);
  dim_force = parse_dim(arg[4]);
  strength = atof(arg[5]);
  bisect_frac = narg > 6 ? atof(arg[6]) : 0.5;
}

int MesoFixPoiseuille::setmask() { return POST_FORCE | MIN_POST_FORCE; }

int MesoFixPoiseuille::register_fix(meso_ctx *ctx)
{
  return meso_fix_pois(ctx,groupbit,dim_ortho,dim_force,strength,bisect_frac);
}
