"""cans_b200 -- B200-native FFT-based Poisson/Helmholtz solver for CaNS.

Host-side mirror of the reference's solver interface (initsolver / fftini /
solver / solve_helmholtz, see cans_b200/solver.py) on top of the C ABI of
include/cans_b200.h.  Importing the package loads the CUDA library and fails
loudly if it has not been built."""
from . import _lib  # noqa: F401  (raises if libcans_b200.so is missing)
from .solver import (Context, Plan, SolverData, initsolver, fftini, fftend, solver, solver_fillps, solve_helmholtz, solver_gaussel_z,  # noqa: F401
                     eigenvalues, tridmatrix, bc_rhs, find_fft, updt_rhs_b, fillps, correc, chkdiv, fill_hash)

__all__ = ["Context", "Plan", "SolverData", "initsolver", "fftini", "fftend", "solver", "solver_fillps", "solve_helmholtz", "solver_gaussel_z",
           "eigenvalues", "tridmatrix", "bc_rhs", "find_fft", "updt_rhs_b", "fillps", "correc", "chkdiv", "fill_hash"]
