/* pair_dpd_fast_meso.h -- pair_style dpd/fast/meso <cut_global> <seed>   (UM/pair_dpd_fast_meso.h, .cu:91-270)
   Same deck grammar and coefficient table as dpd/meso; the pair term is evaluated in fp32 with fp64
   accumulation to memory (A10 in SURVEY.md s8). */
#ifdef PAIR_CLASS

PairStyle(dpd/fast/meso,MesoPairDPDFast)

#else

#ifndef LMP_MESO_PAIR_DPD_FAST
#define LMP_MESO_PAIR_DPD_FAST

#include "pair_dpd_meso.h"

namespace LAMMPS_NS {

class MesoPairDPDFast : public MesoPairDPD {
 public:
  MesoPairDPDFast(class LAMMPS *lmp) : MesoPairDPD(lmp) { precision = MESO_SP; }
};

}

#endif
#endif
