/* Force-included into the UNMODIFIED src/lammps.cpp together with -D__NVCC__ (lammps/Makefile): the core asks
   the CUDA runtime for the device count only when it is compiled by nvcc (src/lammps.cpp:439-441) and divides
   by that count; compiled by g++ it gets the number from the library instead. */
#ifndef LMP_MESO_DEVICE_COUNT_H
#define LMP_MESO_DEVICE_COUNT_H
#include "meso_b200.h"
static inline int cudaGetDeviceCount(int *n) { int c = meso_device_count(); *n = c > 0 ? c : 1; return 0; }
#endif
