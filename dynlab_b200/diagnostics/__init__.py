"""Public diagnostics of dynlab_b200: the seven classes the reference exports
(R:src/dynlab/diagnostics/__init__.py), each backed by the CUDA kernels of ``csrc/``.

    Lagrangian  FTLE, LCS                         K1 (RK45 flow map) + K2 (FTLE stencil)
    Eulerian    AttractionRate, RepulsionRate,    K0 / fused flow evaluation + K3 (stencil +
                TrajectoryRepulsionRate,          rate-of-strain eigen / quadratic forms),
                TrajectoryRepulsionRatio, iLES    ridge field + contour for iLES
"""
from . import eulerian as _eulerian
from . import lagrangian as _lagrangian

FTLE = _lagrangian.FTLE
LCS = _lagrangian.LCS
AttractionRate = _eulerian.AttractionRate
RepulsionRate = _eulerian.RepulsionRate
TrajectoryRepulsionRate = _eulerian.TrajectoryRepulsionRate
TrajectoryRepulsionRatio = _eulerian.TrajectoryRepulsionRatio
iLES = _eulerian.iLES

__all__ = ["FTLE", "LCS", "AttractionRate", "RepulsionRate", "TrajectoryRepulsionRate",
           "TrajectoryRepulsionRatio", "iLES"]
