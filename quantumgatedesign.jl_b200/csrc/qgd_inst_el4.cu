// Sweep kernels for up to 128 levels (4 level rows per lane).
#include "qgd_inst.cuh"
QGD_DEFINE_LAUNCHERS(4)
