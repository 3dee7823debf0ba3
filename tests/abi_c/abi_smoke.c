/* abi_smoke.c -- drives include/qgd_b200.h from plain C (no Python, no ctypes): the way a Julia `ccall` binding or any
 * other host would.  A 2-level Rabi problem (reference src/ProblemConstructors/rabi_oscillator.jl:7-22) with a GRAPE
 * control of 3 amplitudes:
 *   qgd_create -> qgd_set_option -> qgd_eval_forward -> qgd_infidelity_real -> qgd_discrete_adjoint -> qgd_destroy,
 * and the adjoint gradient is checked against central finite differences of the infidelity computed through the same
 * ABI (the reference's own consistency check, test/GradientTests/compare_gradients.jl:47-66, tolerance 1e-9 relative
 * there; 1e-7 here because the step is fixed).
 * Exit code 0 and a line starting with "OK" on success.  Without a CUDA device qgd_create must fail with QGD_ECUDA
 * (there is no CPU fallback): the program then prints "OK (no device ...)" and exits 0, which is what the CPU test asserts. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/qgd_b200.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int rc_ = (call);                                                               \
    if (rc_ != QGD_OK) {                                                            \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, qgd_last_error());        \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

static qgd_matrix_t dense2(const double *a) {
  qgd_matrix_t m;
  memset(&m, 0, sizeof m);
  m.kind = QGD_MAT_DENSE; m.nrows = 2; m.ncols = 2; m.dense = a;
  return m;
}

int main(void) {
  const double zero2[4] = {0, 0, 0, 0};
  const double sym[4] = {0, 1, 1, 0};    /* a + a^T, column-major */
  const double asym[4] = {0, -1, 1, 0};  /* a - a^T: a = [0 1; 0 0] -> [0 1; -1 0] column-major {0,-1,1,0} */
  const double u0[4] = {1, 0, 0, 1}, v0[4] = {0, 0, 0, 0};
  double guard[16];
  memset(guard, 0, sizeof guard);
  qgd_matrix_t symop = dense2(sym), asymop = dense2(asym);
  qgd_control_t ctl;
  memset(&ctl, 0, sizeof ctl);
  ctl.type = QGD_CONTROL_GRAPE; ctl.tf = 1.5; ctl.n_amplitudes = 3;
  qgd_problem_t p;
  memset(&p, 0, sizeof p);
  p.N_tot_levels = 2; p.N_ess_levels = 2; p.N_initial_conditions = 2; p.N_operators = 1;
  p.system_sym = dense2(zero2); p.system_asym = dense2(zero2);
  p.sym_operators = &symop; p.asym_operators = &asymop;
  p.u0 = u0; p.v0 = v0;
  p.guard_subspace_projector.kind = QGD_MAT_DENSE; p.guard_subspace_projector.nrows = 4; p.guard_subspace_projector.ncols = 4;
  p.guard_subspace_projector.dense = guard;
  p.tf = 1.5; p.nsteps = 30; p.gmres_abstol = 1e-14; p.gmres_reltol = 1e-14;
  p.preconditioner = QGD_PRECOND_IDENTITY;
  p.controls = &ctl;
  if (qgd_problem_n_coeff(&p) != 6) { fprintf(stderr, "qgd_problem_n_coeff: expected 6\n"); return 1; }

  qgd_handle_t *h = NULL;
  int rc = qgd_create(&p, -1, &h);
  if (rc == QGD_ECUDA) {
    printf("OK (no device: qgd_create returned QGD_ECUDA, \"%s\")\n", qgd_last_error());
    return 0;
  }
  if (rc != QGD_OK) { fprintf(stderr, "qgd_create failed: %d (%s)\n", rc, qgd_last_error()); return 1; }
  if (qgd_set_option(h, 12345, 1) != QGD_EINVAL) { fprintf(stderr, "unknown option key was accepted\n"); return 1; }
  CHECK(qgd_set_option(h, QGD_OPT_STRICT_MGS, 1));

  /* target: X gate, real-stacked [2N, nic] = vcat(real, imag) */
  const double target[8] = {0, 1, 0, 0, /* column 1: real (0,1), imag (0,0) */ 1, 0, 0, 0};
  double pcof[6] = {0.31, 0.52, 0.44, 0.12, -0.23, 0.07};
  const int order = 6;
  double grad[6], infid = 0, guardpen = 0, final_state[8];
  CHECK(qgd_discrete_adjoint(h, pcof, 1, target, order, 0, grad, &infid, &guardpen, NULL, NULL, NULL, NULL, NULL, NULL));
  CHECK(qgd_eval_forward(h, pcof, 1, order, 1, NULL, final_state, NULL));
  double infid2 = 0;
  CHECK(qgd_infidelity_real(h, final_state, target, 1, &infid2));
  if (fabs(infid - infid2) > 1e-13) { fprintf(stderr, "infidelity mismatch %.17g vs %.17g\n", infid, infid2); return 1; }
  double worst = 0;
  for (int t = 0; t < 6; ++t) {
    const double eps = 1e-5, keep = pcof[t];
    double fp, fm;
    pcof[t] = keep + eps;
    CHECK(qgd_eval_forward(h, pcof, 1, order, 1, NULL, final_state, NULL));
    CHECK(qgd_infidelity_real(h, final_state, target, 1, &fp));
    pcof[t] = keep - eps;
    CHECK(qgd_eval_forward(h, pcof, 1, order, 1, NULL, final_state, NULL));
    CHECK(qgd_infidelity_real(h, final_state, target, 1, &fm));
    pcof[t] = keep;
    const double fd = (fp - fm) / (2 * eps);
    const double err = fabs(fd - grad[t]) / fmax(fabs(fd), 1e-3);
    if (err > worst) worst = err;
  }
  qgd_stats_t st;
  CHECK(qgd_get_stats(h, &st));
  CHECK(qgd_synchronize(h, NULL));
  CHECK(qgd_destroy(h));
  if (worst > 1e-7) { fprintf(stderr, "adjoint gradient vs finite differences: %.3e\n", worst); return 1; }
  printf("OK infidelity %.12f, adjoint gradient vs central differences %.2e, kernel launches of the last call %lld\n", infid, worst,
         (long long)st.kernel_launches);
  return 0;
}
