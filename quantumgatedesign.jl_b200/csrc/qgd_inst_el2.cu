// Sweep kernels for up to 64 levels (2 level rows per lane).
#include "qgd_inst.cuh"
QGD_DEFINE_LAUNCHERS(2)
