// Register-operator sweep kernels for Hermite order 10 (M = 5 Taylor derivatives).
#include "qgd_fast_inst.cuh"
QGD_DEFINE_FAST_LAUNCHERS(5)
