#include "math.h"
#include "stdio.h"
#include "stdlib.h"
#include "string.h"
#include "pair_dpd_meso.h"
#include "engine_meso.h"
#include "atom.h"
#include "comm.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "update.h"

using namespace LAMMPS_NS;

#define EPSILON 1.0e-10

/* ---------------------------------------------------------------------- */

MesoPairDPD::MesoPairDPD(LAMMPS *lmp) : Pair(lmp), MesoBridge(lmp),
  precision(MESO_DP), seed(0), coeff_pushed(false), cut_global(0.0),
  cut(NULL), cut_inv(NULL), a0(NULL), gamma(NULL), sigma(NULL), expw(NULL)
{
  split_flag = 1;                 // bulk / border halves can overlap the halo refresh (src/pair.h:50)
  no_virial_fdotr_compute = 1;    // virial comes from the pairwise sum on the device (newton off)
  single_enable = 1;
  writedata = 0;
}

MesoPairDPD::~MesoPairDPD()
{
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
    memory->destroy(cut);
    memory->destroy(cut_inv);
    memory->destroy(a0);
    memory->destroy(gamma);
    memory->destroy(sigma);
    memory->destroy(expw);
  }
}

void MesoPairDPD::allocate()
{
  allocated = 1;
  int n = atom->ntypes;
  memory->create(setflag,n+1,n+1,"pair:setflag");
  for (int i = 1; i <= n; i++)
    for (int j = i; j <= n; j++) setflag[i][j] = 0;
  memory->create(cutsq,n+1,n+1,"pair:cutsq");
  memory->create(cut,n+1,n+1,"pair:cut");
  memory->create(cut_inv,n+1,n+1,"pair:cut_inv");
  memory->create(a0,n+1,n+1,"pair:a0");
  memory->create(gamma,n+1,n+1,"pair:gamma");
  memory->create(sigma,n+1,n+1,"pair:sigma");
  memory->create(expw,n+1,n+1,"pair:expw");
}

/* ----------------------------------------------------------------------
   force evaluation: the kernels run inside the library
------------------------------------------------------------------------- */

void MesoPairDPD::compute_range(int range, int eflag, int vflag)
{
  if (eflag || vflag) ev_setup(eflag,vflag);
  else evflag = 0;
  if (eflag_atom || vflag_atom)
    error->all(FLERR,"<MESO> per-atom energy/virial of dpd/meso is kept on the device; use global energy/pressure");
  meso_ctx *ctx = mctx("pair_style dpd/meso");
  if (!coeff_pushed) push_coeff();
  MESO_CALL(meso_set_ntimestep(ctx,(int64_t) update->ntimestep));   // seed_now = premix_TEA<64>(seed, ntimestep)
  MESO_CALL(meso_pair_compute(ctx,range,eflag,vflag));
  if (evflag && range != MESO_BULK) tally_from_device(eflag,vflag);
}

void MesoPairDPD::compute(int eflag, int vflag) { compute_range(MESO_LOCAL,eflag,vflag); }
void MesoPairDPD::compute_bulk(int eflag, int vflag) { compute_range(MESO_BULK,eflag,vflag); }
void MesoPairDPD::compute_border(int eflag, int vflag) { compute_range(MESO_BORDER,eflag,vflag); }

void MesoPairDPD::tally_from_device(int eflag, int vflag)
{
  if (!(eflag || vflag)) return;
  if (!evflag) ev_setup(eflag,vflag);
  double v6[6], e = 0.0;
  MESO_CALL(meso_compute_virial(mctx("pair_style dpd/meso"),v6,&e));
  if (eflag_global) eng_vdwl = e;
  if (vflag_global) for (int k = 0; k < 6; k++) virial[k] = v6[k];
}

/* ----------------------------------------------------------------------
   pair_style dpd/meso cut_global seed
------------------------------------------------------------------------- */

void MesoPairDPD::settings(int narg, char **arg)
{
  if (narg != 2) error->all(FLERR,"Illegal pair_style command");

  cut_global = force->numeric(FLERR,arg[0]);
  seed = force->inumeric(FLERR,arg[1]);

  // a new global cutoff replaces the cutoff of every pair that was set explicitly
  if (allocated)
    for (int i = 1; i <= atom->ntypes; i++)
      for (int j = i+1; j <= atom->ntypes; j++)
        if (setflag[i][j]) {
          cut[i][j] = cut_global;
          cut_inv[i][j] = 1.0/cut_global;
        }
  coeff_pushed = false;
}

/* ----------------------------------------------------------------------
   pair_coeff I J a0 gamma sigma expw [cut]
------------------------------------------------------------------------- */

void MesoPairDPD::coeff(int narg, char **arg)
{
  if (narg < 6 || narg > 7) error->all(FLERR,"Incorrect args for pair coefficients");
  if (!allocated) allocate();

  int ilo,ihi,jlo,jhi;
  force->bounds(arg[0],atom->ntypes,ilo,ihi);
  force->bounds(arg[1],atom->ntypes,jlo,jhi);

  const double a0_one = force->numeric(FLERR,arg[2]);
  const double gamma_one = force->numeric(FLERR,arg[3]);
  const double sigma_one = force->numeric(FLERR,arg[4]);
  const double expw_one = force->numeric(FLERR,arg[5]);
  const double cut_one = (narg == 7) ? force->numeric(FLERR,arg[6]) : cut_global;

  int count = 0;
  for (int i = ilo; i <= ihi; i++)
    for (int j = MAX(jlo,i); j <= jhi; j++) {
      a0[i][j] = a0_one;
      gamma[i][j] = gamma_one;
      sigma[i][j] = sigma_one;
      expw[i][j] = expw_one;
      cut[i][j] = cut_one;
      cutsq[i][j] = cut_one*cut_one;
      cut_inv[i][j] = 1.0/cut_one;
      setflag[i][j] = 1;
      count++;
    }
  coeff_pushed = false;

  if (count == 0) error->all(FLERR,"Incorrect args for pair coefficients");
}

/* ---------------------------------------------------------------------- */

void MesoPairDPD::init_style()
{
  // no host neighbor list is requested: the list is built and consumed on the device.
  // Ghost velocities are part of the halo (drag term and RNG signature), newton must be off:
  // ModifiedVerlet::init enforces both.
  mdev("pair_style dpd/meso");
  coeff_pushed = false;
}

double MesoPairDPD::init_one(int i, int j)
{
  if (setflag[i][j] == 0) error->all(FLERR,"All pair coeffs are not set");

  cut[j][i] = cut[i][j];
  cut_inv[j][i] = cut_inv[i][j];
  a0[j][i] = a0[i][j];
  gamma[j][i] = gamma[i][j];
  sigma[j][i] = sigma[i][j];
  expw[j][i] = expw[i][j];
  return cut[i][j];
}

/* ----------------------------------------------------------------------
   coefficient rows for the device: [ (i-1)*ntypes + (j-1) ][7]
------------------------------------------------------------------------- */

void MesoPairDPD::push_coeff()
{
  meso_ctx *ctx = mctx("pair_style dpd/meso");
  const int n = atom->ntypes;
  double *rows = new double[(size_t) n*n*7];
  for (int i = 1; i <= n; i++)
    for (int j = 1; j <= n; j++) {
      const int a = MIN(i,j), b = MAX(i,j);
      if (!setflag[a][b]) { delete [] rows; error->all(FLERR,"All pair coeffs are not set"); }
      double *r = rows + ((size_t) (i-1)*n + (j-1))*7;
      r[0] = cut[a][b];
      r[1] = cut[a][b]*cut[a][b];
      r[2] = cut_inv[a][b];
      r[3] = expw[a][b];
      r[4] = a0[a][b];
      r[5] = gamma[a][b];
      r[6] = sigma[a][b];
    }
  int rc = meso_set_types(ctx,n,atom->mass);
  if (rc == MESO_OK) rc = meso_pair_dpd_settings(ctx,precision,cut_global,seed);
  if (rc == MESO_OK) rc = meso_pair_dpd_coeff(ctx,rows);
  delete [] rows;
  MESO_CALL(rc);
  coeff_pushed = true;
}

/* ----------------------------------------------------------------------
   restart records: per pair {a0,gamma,sigma,expw,cut}, settings {cut_global,seed,mix_flag}
------------------------------------------------------------------------- */

void MesoPairDPD::write_restart(FILE *fp)
{
  write_restart_settings(fp);
  for (int i = 1; i <= atom->ntypes; i++)
    for (int j = i; j <= atom->ntypes; j++) {
      fwrite(&setflag[i][j],sizeof(int),1,fp);
      if (setflag[i][j]) {
        double rec[5] = {a0[i][j],gamma[i][j],sigma[i][j],expw[i][j],cut[i][j]};
        fwrite(rec,sizeof(double),5,fp);
      }
    }
}

void MesoPairDPD::read_restart(FILE *fp)
{
  read_restart_settings(fp);
  allocate();
  const int me = comm->me;
  for (int i = 1; i <= atom->ntypes; i++)
    for (int j = i; j <= atom->ntypes; j++) {
      if (me == 0) fread(&setflag[i][j],sizeof(int),1,fp);
      MPI_Bcast(&setflag[i][j],1,MPI_INT,0,world);
      if (!setflag[i][j]) continue;
      double rec[5];
      if (me == 0) fread(rec,sizeof(double),5,fp);
      MPI_Bcast(rec,5,MPI_DOUBLE,0,world);
      a0[i][j] = rec[0]; gamma[i][j] = rec[1]; sigma[i][j] = rec[2]; expw[i][j] = rec[3]; cut[i][j] = rec[4];
      cutsq[i][j] = cut[i][j]*cut[i][j];
      cut_inv[i][j] = 1.0/cut[i][j];
    }
}

void MesoPairDPD::write_restart_settings(FILE *fp)
{
  fwrite(&cut_global,sizeof(double),1,fp);
  fwrite(&seed,sizeof(int),1,fp);
  fwrite(&mix_flag,sizeof(int),1,fp);
}

void MesoPairDPD::read_restart_settings(FILE *fp)
{
  if (comm->me == 0) {
    fread(&cut_global,sizeof(double),1,fp);
    fread(&seed,sizeof(int),1,fp);
    fread(&mix_flag,sizeof(int),1,fp);
  }
  MPI_Bcast(&cut_global,1,MPI_DOUBLE,0,world);
  MPI_Bcast(&seed,1,MPI_INT,0,world);
  MPI_Bcast(&mix_flag,1,MPI_INT,0,world);
}

/* ----------------------------------------------------------------------
   conservative part only, as a host formula (the drag and random parts need v and the RNG stream)
------------------------------------------------------------------------- */

double MesoPairDPD::single(int i, int j, int itype, int jtype, double rsq, double factor_coul, double factor_dpd,
                           double &fforce)
{
  const double r = sqrt(rsq);
  if (r < EPSILON) {
    fforce = 0.0;
    return 0.5*a0[itype][jtype]*cut[itype][jtype];
  }
  const double w = 1.0 - r*cut_inv[itype][jtype];
  fforce = a0[itype][jtype]*w*factor_dpd/r;
  return factor_dpd*0.5*a0[itype][jtype]*cut[itype][jtype]*w*w;
}
