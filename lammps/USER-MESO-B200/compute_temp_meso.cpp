#include "mpi.h"
#include "compute_styles_meso.h"
#include "engine_meso.h"
#include "atom.h"
#include "domain.h"
#include "error.h"
#include "fix.h"
#include "force.h"
#include "group.h"
#include "modify.h"
#include "update.h"

using namespace LAMMPS_NS;

MesoComputeTemp::MesoComputeTemp(LAMMPS *lmp, int narg, char **arg) : Compute(lmp,narg,arg), MesoBridge(lmp),
  fix_dof(0), tfactor(0.0)
{
  if (narg != 3) error->all(FLERR,"Illegal compute temp command");
  scalar_flag = 1;
  vector_flag = 0;
  extscalar = 0;
  tempflag = 1;
}

void MesoComputeTemp::setup()
{
  fix_dof = 0;
  for (int i = 0; i < modify->nfix; i++) fix_dof += modify->fix[i]->dof(igroup);
  dof_compute();
}

void MesoComputeTemp::dof_compute()
{
  double natoms = group->count(igroup);
  dof = domain->dimension*natoms - (extra_dof + fix_dof);
  tfactor = (dof > 0.0) ? force->mvv2e/(dof*force->boltz) : 0.0;
}

double MesoComputeTemp::compute_scalar()
{
  invoked_scalar = update->ntimestep;
  MesoDevice *dev = mdev("compute temp/meso");
  if (!dev->resident())
    error->all(FLERR,"<MESO> compute temp/meso used before the atoms are on the device (it is valid during a mvv/meso run)");
  double mv2 = 0.0, count = 0.0, all = 0.0;
  MESO_CALL(meso_compute_ke(dev->ctx,groupbit,&mv2,&count));
  MPI_Allreduce(&mv2,&all,1,MPI_DOUBLE,MPI_SUM,world);     // the library returns rank-local sums under MPI
  if (dynamic) dof_compute();
  scalar = all*tfactor;
  return scalar;
}
