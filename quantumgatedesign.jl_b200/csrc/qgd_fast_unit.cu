// Register-operator sweep kernels (qgd_fast.cuh) for ONE Taylor depth M = order / 2, compiled once per depth by the
// Makefile (-DQGD_FAST_M=<m>: objects qgd_fast_m<m>.o; with -DQGD_FAST_STRICT=1 the strict-MGS sweeps, qgd_fast_s_m<m>.o; with
// -DQGD_FAST_FORCED=1 the forced forward solves, qgd_fast_f_m<m>.o; with -DQGD_FAST_TEAM=1 the four-warp latency team,
// qgd_fast_t_m<m>.o; with -DQGD_FAST_RS=1 the row-split groups for 64 < N <= 256, qgd_fast_r_m<m>.o, and with
// -DQGD_FAST_FORCED=1 on top the forced solves on those groups, qgd_fast_rf_m<m>.o)
// so that the heavily unrolled kernels build in parallel.
#ifndef QGD_FAST_M
#error "compile with -DQGD_FAST_M=<order/2>"
#endif
#include "qgd_fast_inst.cuh"
#define QGD_UNIT_EXPAND(MACRO, M) MACRO(M)
#if defined(QGD_FAST_RS) && QGD_FAST_RS && defined(QGD_FAST_FORCED) && QGD_FAST_FORCED
QGD_UNIT_EXPAND(QGD_DEFINE_FAST_LAUNCHERS_RS_FORCED, QGD_FAST_M)
#elif defined(QGD_FAST_RS) && QGD_FAST_RS
QGD_UNIT_EXPAND(QGD_DEFINE_FAST_LAUNCHERS_RS, QGD_FAST_M)
#elif defined(QGD_FAST_TEAM) && QGD_FAST_TEAM
QGD_UNIT_EXPAND(QGD_DEFINE_FAST_LAUNCHERS_TEAM, QGD_FAST_M)
#elif defined(QGD_FAST_FORCED) && QGD_FAST_FORCED
QGD_UNIT_EXPAND(QGD_DEFINE_FAST_LAUNCHERS_FORCED, QGD_FAST_M)
#elif defined(QGD_FAST_STRICT) && QGD_FAST_STRICT
QGD_UNIT_EXPAND(QGD_DEFINE_FAST_LAUNCHERS_STRICT, QGD_FAST_M)
#else
QGD_UNIT_EXPAND(QGD_DEFINE_FAST_LAUNCHERS, QGD_FAST_M)
#endif
