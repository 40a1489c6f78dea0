/* pair_dpd_meso.h -- pair_style dpd/meso <cut_global> <seed>   (UM/pair_dpd_meso.h, UM/pair_dpd_meso.cu:241-467)
   pair_coeff I J a0 gamma sigma expw [cut]
   fp64 arithmetic on the fp32-packed coordinates (A11 in SURVEY.md s8); the force kernel, the per-pair
   TEA Gaussian and the neighbor table live in the library (meso_pair_compute).  This class parses the
   deck, keeps the per-type-pair tables LAMMPS expects (cutsq, setflag, restart records) and hands the
   7-column coefficient rows {cut,cutsq,cutinv,expw,a0,gamma,sigma} (UM/pair_dpd_meso.h:15-24) to the library. */
#ifdef PAIR_CLASS

PairStyle(dpd/meso,MesoPairDPD)

#else

#ifndef LMP_MESO_PAIR_DPD
#define LMP_MESO_PAIR_DPD

#include "pair.h"
#include "meso_bridge.h"

namespace LAMMPS_NS {

class MesoPairDPD : public Pair, protected MesoBridge {
 public:
  MesoPairDPD(class LAMMPS *);
  virtual ~MesoPairDPD();
  virtual void compute(int, int);
  virtual void compute_bulk(int, int);
  virtual void compute_border(int, int);
  virtual void settings(int, char **);
  virtual void coeff(int, char **);
  virtual void init_style();
  virtual double init_one(int, int);
  virtual void write_restart(FILE *);
  virtual void read_restart(FILE *);
  virtual void write_restart_settings(FILE *);
  virtual void read_restart_settings(FILE *);
  virtual double single(int, int, int, int, double, double, double, double &);

  // called by the integrator
  void push_coeff();                 // prepare_coeff, UM/pair_dpd_meso.cu:68-89
  void tally_from_device(int, int);  // global energy / virial of the last force evaluation -> eng_vdwl, virial[6]

 protected:
  int precision;                     // MESO_DP here, MESO_SP in dpd/fast/meso
  int seed;
  bool coeff_pushed;
  double cut_global;
  double **cut, **cut_inv, **a0, **gamma, **sigma, **expw;

  void allocate();
  void compute_range(int, int, int);
};

}

#endif
#endif
