#include "string.h"
#include "fix_styles_meso.h"
#include "error.h"
#include "update.h"

using namespace LAMMPS_NS;

void MesoFixResident::init()
{
  if (strstr(update->integrate_style,"meso") == NULL)
    error->all(FLERR,"<MESO> device-resident fixes need run_style mvv/meso (or verlet/meso)");
  handle = register_fix(mctx(style));
  if (handle < 0) MESO_CALL(handle);
}

void MesoFixResident::post_force(int vflag) { MESO_CALL(meso_fix_post_force(mctx(style),handle)); }

void MesoFixResident::bounce() { MESO_CALL(meso_fix_bounce(mctx(style),handle)); }
