// Sweep kernels for up to 256 levels (8 level rows per lane): the dense 4-qudit x 4-level shape.
#include "qgd_inst.cuh"
QGD_DEFINE_LAUNCHERS(8)
