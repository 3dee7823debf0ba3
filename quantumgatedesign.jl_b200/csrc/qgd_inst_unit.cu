// Generic row-ELL sweep kernels (qgd_kernels.cuh) for ONE lane width EL = level rows per lane (N <= 32 EL), compiled once
// per width by the Makefile (-DQGD_EL=<el>: objects qgd_inst_el<el>.o).
#ifndef QGD_EL
#error "compile with -DQGD_EL=<1|2|4|8>"
#endif
#include "qgd_inst.cuh"
#define QGD_UNIT_EXPAND(MACRO, EL) MACRO(EL)
QGD_UNIT_EXPAND(QGD_DEFINE_LAUNCHERS, QGD_EL)
