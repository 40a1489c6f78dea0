#include "string.h"
#include "bond_harmonic_meso.h"
#include "pair_dpd_meso.h"
#include "atom.h"
#include "error.h"
#include "force.h"
#include "update.h"

using namespace LAMMPS_NS;

void MesoBondHarmonic::init_style()
{
  if (strstr(update->integrate_style,"meso") == NULL)
    error->all(FLERR,"<MESO> bond_style harmonic/meso needs run_style mvv/meso (or verlet/meso)");
  push_coeff();
}

void MesoBondHarmonic::push_coeff()
{
  MESO_CALL(meso_bond_harmonic_coeff(mctx("bond_style harmonic/meso"),atom->nbondtypes,k,r0));
}

void MesoBondHarmonic::compute(int eflag, int vflag)
{
  if (eflag || vflag) ev_setup(eflag,vflag);
  else evflag = 0;
  MESO_CALL(meso_bond_compute(mctx("bond_style harmonic/meso"),eflag,vflag));
  tally_from_device(eflag,vflag);
}

/* the bond kernel adds its per-atom virial to the array the pair kernel filled, so the global virial is handed out
   once, through the pair style (it is re-read here because Pair::compute tallied before the bonds ran) */
void MesoBondHarmonic::tally_from_device(int eflag, int vflag)
{
  if (!(eflag || vflag)) return;
  if (eflag) {
    double e = 0.0;
    MESO_CALL(meso_compute_bond_energy(mctx("bond_style harmonic/meso"),&e));
    energy = e;
  }
  MesoPairDPD *dpd = dynamic_cast<MesoPairDPD *>(force->pair);
  if (vflag && dpd) dpd->tally_from_device(eflag,vflag);
}
