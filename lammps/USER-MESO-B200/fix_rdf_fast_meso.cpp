#include "math.h"
#include "stdio.h"
#include "stdlib.h"
#include "string.h"
#include "fix_styles_meso.h"
#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "group.h"
#include "pair.h"

using namespace LAMMPS_NS;
using namespace FixConst;

/* argument grammar and messages of UM/fix_rdf_fast_meso.cu:42-80 */
MesoFixRDFFast::MesoFixRDFFast(LAMMPS *lmp, int narg, char **arg) :
  MesoFixResident(lmp,narg,arg), n_bin(0), n_every(1), rc(0.0), n_i(0.0), n_j(0.0), n_steps(0.0), last_steps(0.0)
{
  j_groupbit = groupbit;
  for (int i = 0; i < narg; i++) {
    if (!strcmp(arg[i],"output")) {
      if (++i >= narg) error->all(FLERR,"Incomplete compute vprof command after 'output'");
      output = arg[i];
    } else if (!strcmp(arg[i],"nbin")) {
      if (++i >= narg) error->all(FLERR,"Incomplete compute vprof command after 'nbin'");
      n_bin = atoi(arg[i]);
    } else if (!strcmp(arg[i],"every")) {
      if (++i >= narg) error->all(FLERR,"Incomplete compute vprof command after 'every'");
      n_every = atoi(arg[i]);
    } else if (!strcmp(arg[i],"other")) {
      if (++i >= narg) error->all(FLERR,"Incomplete compute vprof command after 'other'");
      int j_group = group->find(arg[i]);
      if (j_group == -1) error->all(FLERR,"<MESO> Undefined other group id in fix rdf/meso");
      j_groupbit = group->bitmask[j_group];
    }
  }
  if (output == "" || n_bin == 0) error->all(FLERR,"Incomplete compute rdf command: insufficient arguments");
  if (!force->pair) error->all(FLERR,"<MESO> fix rho/meso must be used together with a pair style");
  total.assign(n_bin,0.0);
  last.assign(n_bin,0.0);
}

MesoFixRDFFast::~MesoFixRDFFast() { dump(); }

int MesoFixRDFFast::setmask() { return POST_FORCE; }

void MesoFixRDFFast::init()
{
  rc = force->pair->cutforce;
  MesoFixResident::init();               // a fresh registration starts from an empty device histogram
  last.assign(n_bin,0.0);
  last_steps = 0.0;
}

int MesoFixRDFFast::register_fix(meso_ctx *ctx) { return meso_fix_rdf(ctx,groupbit,j_groupbit,n_bin,n_every); }

/* the device histogram lives as long as the registration; fold what was added since the last harvest into the host totals */
void MesoFixRDFFast::harvest()
{
  if (handle < 0) return;
  std::vector<double> now(n_bin,0.0);
  double steps = 0.0, ni = 0.0, nj = 0.0;
  MESO_CALL(meso_fix_rdf_read(mctx(style),handle,n_bin,&now[0],&steps,&ni,&nj));
  for (int b = 0; b < n_bin; b++) { total[b] += now[b] - last[b]; last[b] = now[b]; }
  n_steps += steps - last_steps;
  last_steps = steps;
  n_i = ni; n_j = nj;                    // this rank's part; dump() sums over the ranks
}

void MesoFixRDFFast::post_run() { harvest(); }

/* MesoFixRDFFast::dump, UM/fix_rdf_fast_meso.cu:182-219 (pi = 3.1415 as there) */
void MesoFixRDFFast::dump()
{
  double n[2] = {n_i, n_j}, nall[2];
  MPI_Allreduce(n,nall,2,MPI_DOUBLE,MPI_SUM,world);
  std::vector<double> master(n_bin,0.0);
  MPI_Allreduce(&total[0],&master[0],n_bin,MPI_DOUBLE,MPI_SUM,world);
  double V = 1.0;
  for (int d = 0; d < 3; d++) V *= domain->boxhi[d] - domain->boxlo[d];
  const double rho = nall[1]/V;
  if (comm->me == 0 && n_steps > 0.0 && nall[0] > 0.0 && rho > 0.0) {
    FILE *fp = fopen(output.c_str(),"w");
    if (fp) {
      const double bin_sz = rc/n_bin;
      for (int i = 0; i < n_bin; i++) {
        const double freq = master[i]/nall[0]/n_steps;
        const double gr = freq/(4./3.*3.1415*(pow(bin_sz*(i+1),3.) - pow(bin_sz*i,3.)))/rho;
        fprintf(fp,"%.15g\t%.15g\t\n",(i+0.5)*bin_sz,gr);
      }
      fclose(fp);
    }
  }
}
