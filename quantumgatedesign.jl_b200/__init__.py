"""quantumgatedesign.jl_b200 -- B200-native gradient hot path of QuantumGateDesign.jl.

Host-side mirror of the reference's API for the path (SchrodingerProb, the control types,
eval_forward, discrete_adjoint, infidelity ...) over the C ABI of csrc/libqgd_b200.so
(include/qgd_b200.h).  The directory name contains a dot, so the package is loaded through
`__graft_entry__.load_package()` under the module name `qgd_b200`.
"""
from . import _abi
from . import configs
from .controls import (AbstractControl, BSpline2Control, CarrierControl, FortranBSplineControl, GRAPEControl,
                       HostEvaluatedControl, SinCosControl, as_control_list, build_control_tables, control_slices,
                       get_number_of_control_parameters, has_host_controls)
from .problem import (DiagonalHamiltonianPreconditioner, DispersiveProblem, IdentityPreconditioner, LUPreconditioner,
                      SchrodingerProb, complex_to_real, construct_rabi_prob, construct_rand_prob, control_ops,
                      create_gate, create_initial_conditions, guard_projector, lowering_operators_system,
                      multi_qudit_hamiltonian_dispersive, real_to_complex)
from . import backend
from . import distributed
from .api import (discrete_adjoint, discrete_adjoint_, discrete_adjoint_batch, eval_forward, eval_forward_,
                  eval_grad_forced, guard_penalty_real, infidelity, infidelity_real)
from .backend import Handle, MultiGPU, QGDError, comm_unique_id, get_handle, measure_dmma_peak, measure_fp64_peak
from .optimize import optimize_gate
from .convergence import (estimate_N_timesteps, estimate_timesteps_per_period, get_histories, get_shortest_period,
                          richardson_extrap_rel_err, richardson_extrap_sol)
