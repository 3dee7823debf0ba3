// Sweep kernels for up to 32 levels (1 level rows per lane).
#include "qgd_inst.cuh"
QGD_DEFINE_LAUNCHERS(1)
