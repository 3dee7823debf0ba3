// Register-operator sweep kernels for Hermite order 8 (M = 4 Taylor derivatives).
#include "qgd_fast_inst.cuh"
QGD_DEFINE_FAST_LAUNCHERS(4)
