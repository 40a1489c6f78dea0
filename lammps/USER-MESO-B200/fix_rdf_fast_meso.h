/* fix_rdf_fast_meso.h -- fix ID group rdf/fast/meso output <file> nbin <n> [every <k>] [other <group>]
   (UM/fix_rdf_fast_meso.h, UM/fix_rdf_fast_meso.cu:42-219): radial distribution function up to the pair cutoff from the
   device neighbor table, sampled in the post_force slot; g(r) is written when the fix is destroyed. */
#ifdef FIX_CLASS

FixStyle(rdf/fast/meso,MesoFixRDFFast)

#else

#ifndef LMP_MESO_FIX_RDF_FAST
#define LMP_MESO_FIX_RDF_FAST

#include <string>
#include <vector>
#include "fix_resident_meso.h"

namespace LAMMPS_NS {

class MesoFixRDFFast : public MesoFixResident {
 public:
  MesoFixRDFFast(class LAMMPS *, int, char **);
  virtual ~MesoFixRDFFast();
  virtual int setmask();
  virtual void init();
  virtual void post_run();
 protected:
  std::string output;
  int n_bin, n_every, j_groupbit;
  double rc, n_i, n_j, n_steps;
  std::vector<double> total, last;       // pair counts per bin: harvested from the device / at the previous harvest
  double last_steps;
  virtual int register_fix(meso_ctx *);
  void harvest();
  void dump();
};

}

#endif
#endif
