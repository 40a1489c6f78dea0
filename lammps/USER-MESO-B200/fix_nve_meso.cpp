#include "string.h"
#include "fix_styles_meso.h"
#include "error.h"
#include "update.h"

using namespace LAMMPS_NS;
using namespace FixConst;

FixNVEMeso::FixNVEMeso(LAMMPS *lmp, int narg, char **arg) : Fix(lmp,narg,arg), MesoBridge(lmp)
{
  if (narg < 3) error->all(FLERR,"Illegal fix nve/meso command");
  time_integrate = 1;
}

int FixNVEMeso::setmask() { return INITIAL_INTEGRATE | FINAL_INTEGRATE; }

void FixNVEMeso::init()
{
  if (strstr(update->integrate_style,"meso") == NULL)
    error->all(FLERR,"<MESO> fix nve/meso needs run_style mvv/meso (or verlet/meso)");
  reset_dt();
}

void FixNVEMeso::reset_dt() { MESO_CALL(meso_set_timestep_size(mctx("fix nve/meso"),update->dt)); }

void FixNVEMeso::initial_integrate(int vflag) { MESO_CALL(meso_initial_integrate(mctx("fix nve/meso"),groupbit)); }

void FixNVEMeso::final_integrate() { MESO_CALL(meso_final_integrate(mctx("fix nve/meso"),groupbit)); }
