#include "string.h"
#include "run_style_meso.h"
#include "engine_meso.h"
#include "fix_styles_meso.h"
#include "bond_harmonic_meso.h"
#include "pair_dpd_meso.h"
#include "atom.h"
#include "comm.h"
#include "compute.h"
#include "domain.h"
#include "error.h"
#include "fix.h"
#include "force.h"
#include "modify.h"
#include "neighbor.h"
#include "output.h"
#include "pair.h"
#include "timer.h"
#include "update.h"

using namespace LAMMPS_NS;

/* ---------------------------------------------------------------------- */

ModifiedVerlet::ModifiedVerlet(LAMMPS *lmp, int narg, char **arg) :
  Integrate(lmp,narg,arg), MesoBridge(lmp), fused_groupbit(-1), dpd(NULL) {}

/* ----------------------------------------------------------------------
   initialization before run: the compliance rules of UM/mvv_meso.cu:79-133
------------------------------------------------------------------------- */

void ModifiedVerlet::init()
{
  Integrate::init();
  mdev("run_style mvv/meso");

  if (modify->nfix == 0 && comm->me == 0) error->warning(FLERR,"No fixes defined, atoms won't move");

  if (force->newton_pair || force->newton || force->newton_bond) {
    if (comm->me == 0)
      error->warning(FLERR,"<MESO> newton_pair not allowed in MESO mode, forced to 0; ghost_velocity forced to 1");
    force->newton = force->newton_pair = force->newton_bond = 0;
  }
  comm->ghost_velocity = 1;
  virial_style = 1;                              // explicit pairwise virial: there is no reverse communication
  ev_setup();

  if (domain->triclinic) error->one(FLERR,"<MESO> triclinic domain not supported in USER-MESO");
  if (force->kspace) error->one(FLERR,"<MESO> kspace not supported in USER-MESO");
  if (atom->sortfreq > 0) {
    if (comm->me == 0) error->warning(FLERR,"<MESO> atom sort frequency managed automatically in USER-MESO");
    atom->sortfreq = 0;
  }

  dpd = dynamic_cast<MesoPairDPD *>(force->pair);
  if (dpd == NULL) error->all(FLERR,"<MESO> run_style mvv/meso needs pair_style dpd/meso or dpd/fast/meso");
  if (force->bond && dynamic_cast<MesoBondHarmonic *>(force->bond) == NULL)
    error->all(FLERR,"<MESO> bond styles other than harmonic/meso are not part of USER-MESO-B200");
  if (force->angle || force->dihedral || force->improper)
    error->all(FLERR,"<MESO> angle, dihedral and improper styles are not part of USER-MESO-B200");

  // every fix acts on device-resident atoms, so it has to be a /meso style;
  // exactly one nve/meso plus fixes that live in the library's own fix list => the fused run loop.
  // The list is rebuilt by the fixes' init() (Modify::init runs after this function).
  MESO_CALL(meso_fix_clear(mctx("run_style mvv/meso")));
  int n_nve = 0, others = 0, gbit = -1;
  for (int i = 0; i < modify->nfix; i++) {
    Fix *fix = modify->fix[i];
    const char *s = fix->style;
    const int len = strlen(s);
    if (len < 5 || strcmp(s+len-5,"/meso") != 0) {
      char msg[256];
      sprintf(msg,"<MESO> fix style %.100s does not act on device-resident atoms (use a /meso style)",s);
      error->all(FLERR,msg);
    }
    FixNVEMeso *nve = dynamic_cast<FixNVEMeso *>(fix);
    if (nve) { n_nve++; gbit = nve->group_bit(); }
    else if (dynamic_cast<MesoFixResident *>(fix) == NULL) others++;
  }
  fused_groupbit = (n_nve == 1 && others == 0) ? gbit : -1;
}

/* ----------------------------------------------------------------------
   setup before run (UM/mvv_meso.cu:139-219): host arrays are authoritative between runs
------------------------------------------------------------------------- */

void ModifiedVerlet::device_setup(int outflag)
{
  MesoDevice *dev = mdev("run_style mvv/meso");
  update->setupflag = 1;

  atom->setup();
  modify->setup_pre_exchange();
  domain->pbc();
  domain->reset_box();
  comm->setup();

  dev->push_settings();
  dpd->push_coeff();
  MesoBondHarmonic *bond = dynamic_cast<MesoBondHarmonic *>(force->bond);
  if (bond) bond->push_coeff();
  dev->upload_atoms();

  // wrap + reorder + ghosts + neighbor table + forces of step `ntimestep`, all on the device
  ev_set(update->ntimestep);
  MESO_CALL(meso_setup(dev->ctx,eflag,vflag));
  dpd->tally_from_device(eflag,vflag);
  if (bond) bond->tally_from_device(eflag,vflag);

  modify->setup(vflag);
  if (outflag) {
    dev->download_atoms();
    output->setup();
  }
  update->setupflag = 0;
}

void ModifiedVerlet::setup()
{
  if (comm->me == 0 && screen) fprintf(screen,"Setting up run ...\n");
  device_setup(1);
}

void ModifiedVerlet::setup_minimal(int flag)
{
  MesoDevice *dev = mdev("run_style mvv/meso");
  if (flag || !dev->resident()) { device_setup(0); return; }
  update->setupflag = 1;
  ev_set(update->ntimestep);
  force_clear();
  force->pair->compute(eflag,vflag);
  if (force->bond) force->bond->compute(eflag,vflag);
  MESO_CALL(meso_fix_post_force(dev->ctx,-1));      // Fix::setup -> post_force of the device-resident fixes
  modify->setup(vflag);
  update->setupflag = 0;
}

/* ----------------------------------------------------------------------
   one step through the phase entry points (any /meso fix list; energy/virial tallies)
------------------------------------------------------------------------- */

void ModifiedVerlet::force_clear()
{
  MESO_CALL(meso_force_clear(mctx("run_style mvv/meso"),MESO_LOCAL,(eflag || vflag) ? 1 : 0));
}

void ModifiedVerlet::step_by_phases(bigint ntimestep)
{
  meso_ctx *ctx = mctx("run_style mvv/meso");
  MESO_CALL(meso_set_ntimestep(ctx,(int64_t) ntimestep));

  modify->initial_integrate(vflag);
  if (modify->n_post_integrate) modify->post_integrate();

  if (meso_neighbor_decide(ctx)) {
    if (modify->n_pre_exchange) modify->pre_exchange();
    if (modify->n_pre_neighbor) modify->pre_neighbor();
    MESO_CALL(meso_rebuild(ctx));
  } else MESO_CALL(meso_forward_comm(ctx));

  force_clear();
  if (modify->n_pre_force) modify->pre_force(vflag);
  force->pair->compute(eflag,vflag);
  if (force->bond) force->bond->compute(eflag,vflag);          // UM/mvv_meso.cu:385-393
  if (modify->n_post_force) modify->post_force(vflag);
  modify->final_integrate();
  if (modify->n_end_of_step) modify->end_of_step();
}

/* ----------------------------------------------------------------------
   true if every compute the previous thermo output evaluated keeps its data on the device (the /meso computes) or
   only combines host scalars that the device styles tally (stock pressure and pe: Pair::virial, Pair::eng_vdwl)
------------------------------------------------------------------------- */

bool ModifiedVerlet::thermo_reads_device_only()
{
  if (modify->n_end_of_step || output->thermo == NULL) return false;
  const bigint last = output->last_thermo;
  for (int i = 0; i < modify->ncompute; i++) {
    Compute *c = modify->compute[i];
    if (c->invoked_scalar < last && c->invoked_vector < last && c->invoked_array < last &&
        c->invoked_peratom < last && c->invoked_local < last) continue;          // not part of the thermo line
    if (strstr(c->style,"/meso") || strcmp(c->style,"pressure") == 0 || strcmp(c->style,"pe") == 0) continue;
    return false;
  }
  return true;
}

void ModifiedVerlet::flush(int &pending)
{
  if (pending == 0) return;
  MESO_CALL(meso_run(mctx("run_style mvv/meso"),pending,fused_groupbit));
  pending = 0;
}

/* ----------------------------------------------------------------------
   run for N steps
------------------------------------------------------------------------- */

void ModifiedVerlet::run(int n)
{
  MesoDevice *dev = mdev("run_style mvv/meso");
  meso_ctx *ctx = dev->ctx;
  dev->timers_begin();
  dev->profile_run_begin(update->ntimestep);
  MESO_CALL(meso_set_ntimestep(ctx,(int64_t) update->ntimestep));

  int pending = 0;                       // steps queued for one fused meso_run call
  for (int i = 0; i < n; i++) {
    const bigint ntimestep = ++update->ntimestep;
    ev_set(ntimestep);

    if (fused_groupbit >= 0 && !eflag && !vflag) pending++;
    else {
      flush(pending);
      step_by_phases(ntimestep);
    }

    if (ntimestep == output->next) {
      flush(pending);
      // transfer_pre_output (UM/mvv_meso.cu:411-416) copies every per-atom array back at every output step.  Dumps and
      // restarts read host arrays; a thermo line whose computes all reduce on the device does not: no PCIe traffic then
      if (ntimestep == output->next_dump_any || (output->restart_flag && ntimestep == output->next_restart) ||
          !thermo_reads_device_only())
        dev->download_atoms();
      timer->stamp();
      output->write(ntimestep);
      timer->stamp(TIME_OUTPUT);
      dev->profile_step(ntimestep);
    }
  }
  flush(pending);
  MESO_CALL(meso_sync(ctx));
  dev->profile_run_end();
  dev->timers_end();
}

/* ---------------------------------------------------------------------- */

void ModifiedVerlet::cleanup()
{
  // hand the atoms back: the host arrays are what every other LAMMPS command reads
  mdev("run_style mvv/meso")->download_atoms();
  modify->post_run();
  domain->box_too_small_check();
  update->update_time();
}

void ModifiedVerlet::reset_dt()
{
  MESO_CALL(meso_set_timestep_size(mctx("run_style mvv/meso"),update->dt));
}
