// Register-operator sweep kernels for Hermite order 4 (M = 2 Taylor derivatives).
#include "qgd_fast_inst.cuh"
QGD_DEFINE_FAST_LAUNCHERS(2)
