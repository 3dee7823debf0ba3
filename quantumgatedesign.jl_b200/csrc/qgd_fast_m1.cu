// Register-operator sweep kernels for Hermite order 2 (M = 1 Taylor derivatives).
#include "qgd_fast_inst.cuh"
QGD_DEFINE_FAST_LAUNCHERS(1)
