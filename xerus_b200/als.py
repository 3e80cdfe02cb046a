"""Host mirror of the reference's ALSVariant (include/xerus/algorithms/als.h:37-223, src/xerus/algorithms/als.cpp):
the presets ALS, ALS_SPD, DMRG, DMRG_SPD and the call signatures `variant(A, x, b, numHalfSweeps | convergenceEpsilon)`
and `variant(x, b, ...)`.  One call = one device-resident run of ALSVariant::solve (als.cpp:483-553).
"""
import ctypes as C

from . import _lib
from ._lib import XerusError, call
from .tt import TTOperator, TTTensor


class ALSVariant:
    def __init__(self, sites, numHalfSweeps, assumeSPD, localTolerance=0.0, localMaxIterations=0, localSolver="lapack"):
        if sites <= 0:
            raise XerusError(1, "sites must be positive")          # als.h:141
        self.sites = int(sites)
        self.numHalfSweeps = int(numHalfSweeps)
        self.convergenceEpsilon = 1e-6                             # als.h:137
        self.preserveCorePosition = True
        self.assumeSPD = bool(assumeSPD)
        self.localTolerance = float(localTolerance)
        self.localMaxIterations = int(localMaxIterations)
        if localSolver not in ("lapack", "ASD"):
            raise XerusError(1, "localSolver must be 'lapack' (ALSVariant::lapack_solver) or 'ASD' (ALSVariant::ASD_solver)")
        self.localSolver = localSolver                             # als.h:123-124: the reference takes a std::function here
        self.last_local_iterations = 0

    def __call__(self, *args):
        """(A, x, b[, numHalfSweeps:int | convergenceEpsilon:float]) or (x, b[, ...]) — als.h:152-207."""
        args = list(args)
        A = args.pop(0) if isinstance(args[0], TTOperator) else None
        x, b = args[0], args[1]
        rest = args[2:]
        num_half_sweeps, conv = self.numHalfSweeps, self.convergenceEpsilon
        if rest:
            if isinstance(rest[0], float):
                conv = rest[0]
            else:
                num_half_sweeps = int(rest[0])
        if not isinstance(x, TTTensor) or not isinstance(b, TTTensor):
            raise XerusError(1, "x and b must be TTTensors")
        if x.dimensions != b.dimensions:
            raise XerusError(1, "x.dimensions != b.dimensions")    # als.cpp:489
        opt = _lib.ALSOptions()
        call("xb_als_default_options", C.byref(opt), self.sites, int(self.assumeSPD))
        opt.num_half_sweeps = num_half_sweeps
        opt.convergence_epsilon = conv
        opt.preserve_core_position = int(self.preserveCorePosition)
        opt.local_tolerance = self.localTolerance
        opt.local_max_iterations = self.localMaxIterations
        opt.local_solver = 1 if self.localSolver == "ASD" else 0
        energy = C.c_double()
        iters = C.c_size_t()
        call("xb_als_solve", A._h if A is not None else None, x._h, b._h, C.byref(opt), C.byref(energy), C.byref(iters))
        self.last_local_iterations = iters.value
        return energy.value


ALS = ALSVariant(1, 0, False)        # als.cpp:556-563
ALS_SPD = ALSVariant(1, 0, True)
DMRG = ALSVariant(2, 0, False)
DMRG_SPD = ALSVariant(2, 0, True)
ASD = ALSVariant(1, 0, False, localSolver="ASD")
ASD_SPD = ALSVariant(1, 0, True, localSolver="ASD")
