"""nls-b200: B200-native (sm_100a) Newton hot path of NonlinearSolids.jl.

The product is libnls_b200.so (CUDA kernels behind a flat C ABI, include/nls_b200.h). This
package is the host-side mirror of the reference's Julia API for that path -- meshes,
materials, Residual/Jacobian functors, NewtonSolver, PseudoTimeStepper -- calling the ABI
through ctypes exactly as the Julia `ccall` shim in julia/NonlinearSolidsB200.jl does.
"""
from . import _lib, distributed
from .arclength import ArcLength, ArcLengthResult, newtonarclength
from ._lib import Handle, NLSError, lib
from .fem import (Dirichlet, Element, ElementBoundary, FEMModel, Lagrange, Mesh, N, Neumann, Quadrature,
                  TimeDependent, dN, gauss_quadrature, getstate, integrate, interpolate, lagrange)
from .gallery import (CSRMatrix, DistributedForce, Jacobian, Residual, applymass, assemblemass, compute_dinternalforce,
                      compute_internalforce, residual_jacobian)
from .materials import (ExponentialMaterial, J2Plasticity, LinearElasticity1D, LinearElastoPlasticity1D, NeoHookean,
                        initialize_state, save_state)
from .meshes import bar_mesh, face_nodes, grid_mesh, slab_local_mesh, slab_planes, smooth_field
from .solvers import (REASON_CONVERGED_ABSOLUTE, REASON_CONVERGED_RELATIVE, REASON_DIVERGED_LINEAR_SOLVE,
                      REASON_DIVERGED_MAXITS, REASON_DIVERGED_STEP, REASON_NOT_YET_CONVERGED, NewtonSolver,
                      NewmarkTimeStepper, PseudoTimeStepper, is_converged, set_ctx, set_jacobian_fn, set_residual_fn)
