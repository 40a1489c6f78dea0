#include "mpi.h"
#include "string.h"
#include "compute_styles_meso.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "modify.h"
#include "pair.h"
#include "bond.h"
#include "update.h"

using namespace LAMMPS_NS;

MesoComputePE::MesoComputePE(LAMMPS *lmp, int narg, char **arg) : Compute(lmp,narg,arg)
{
  if (narg < 3) error->all(FLERR,"Illegal compute pe command");
  if (igroup) error->all(FLERR,"Compute pe must use group all");
  scalar_flag = 1;
  extscalar = 1;
  peflag = 1;
  timeflag = 1;
  pairflag = bondflag = thermoflag = 1;
  if (narg > 3) {
    pairflag = bondflag = thermoflag = 0;
    for (int iarg = 3; iarg < narg; iarg++) {
      if (strcmp(arg[iarg],"pair") == 0) pairflag = 1;
      else if (strcmp(arg[iarg],"bond") == 0) bondflag = 1;
      else if (strcmp(arg[iarg],"thermo") == 0) thermoflag = 1;
      else error->all(FLERR,"<MESO> compute pe/meso knows the keywords pair, bond and thermo");
    }
  }
}

double MesoComputePE::compute_scalar()
{
  invoked_scalar = update->ntimestep;
  if (update->eflag_global != invoked_scalar) error->all(FLERR,"Energy was not tallied on needed timestep");

  // the pair style filled eng_vdwl from the device reduction (MesoPairDPD::tally_from_device)
  double one = 0.0;
  if (pairflag && force->pair) one += force->pair->eng_vdwl + force->pair->eng_coul;
  if (bondflag && force->bond) one += force->bond->energy;
  MPI_Allreduce(&one,&scalar,1,MPI_DOUBLE,MPI_SUM,world);

  if (pairflag && force->pair && force->pair->tail_flag)
    scalar += force->pair->etail/(domain->xprd*domain->yprd*domain->zprd);
  if (thermoflag && modify->n_thermo_energy) scalar += modify->thermo_energy();
  return scalar;
}
