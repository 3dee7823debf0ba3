// Register-operator sweep kernels for Hermite order 12 (M = 6 Taylor derivatives).
#include "qgd_fast_inst.cuh"
QGD_DEFINE_FAST_LAUNCHERS(6)
