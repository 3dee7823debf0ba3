// Register-operator sweep kernels for Hermite order 6 (M = 3 Taylor derivatives).
#include "qgd_fast_inst.cuh"
QGD_DEFINE_FAST_LAUNCHERS(3)
