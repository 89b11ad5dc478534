"""Eulerian diagnostics (mirror of R:src/dynlab/diagnostics/eulerian.py): AttractionRate,
RepulsionRate, TrajectoryRepulsionRate, TrajectoryRepulsionRatio, iLES — all one fused K3
launch (np.gradient stencil + rate-of-strain tensor + closed-form 2x2 eigen / quadratic form).
"""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from .._engine import Engine
from ._base_classes import EulerianDiagnostic2D, RidgeExtractor2D


class _AttractionRepulsionRate(EulerianDiagnostic2D):
    """s_1 / s_n: smallest / largest eigenvalue of S = (grad v + grad v^T)/2
    (R:eulerian.py:9-61)."""

    def compute(self, x, y, u=None, v=None, f=None, t=None, edge_order: int = 1,
                eigenvalue_index: int = 0):
        field = self._eigen_field_device(x, y, u, v, f, t, edge_order, eigenvalue_index)
        return np.ma.MaskedArray(Engine.get().to_host(field))

    def _eigen_field_device(self, x, y, u, v, f, t, edge_order, eigenvalue_index):
        ud, vd = EulerianDiagnostic2D.compute(self, x, y, u, v, f, t, edge_order)
        if eigenvalue_index not in (0, -1, 1, -2):
            raise IndexError(f"index {eigenvalue_index} is out of bounds for axis 0 with size 2")
        mode = nat.EULER_S_MIN if eigenvalue_index in (0, -2) else nat.EULER_S_MAX
        field, _, _ = self._euler(mode, ud, vd, edge_order)
        self.field_device = field
        return field


class AttractionRate(_AttractionRepulsionRate):
    """Attraction rate s_1 (R:eulerian.py:64-91)."""

    def compute(self, x, y, u=None, v=None, f=None, t=None, edge_order: int = 1):
        self.field = super().compute(x, y, u, v, f, t, edge_order, 0)
        return self.field

    def compute_device(self, x, y, u=None, v=None, f=None, t=None, edge_order: int = 1):
        """Same computation with the result left on the device (``torch.Tensor[ydim, xdim]``);
        ``u, v`` may be CUDA tensors, in which case nothing crosses PCIe."""
        return self._eigen_field_device(x, y, u, v, f, t, edge_order, 0)


class RepulsionRate(_AttractionRepulsionRate):
    """Repulsion rate s_n (R:eulerian.py:94-121)."""

    def compute(self, x, y, u=None, v=None, f=None, t=None, edge_order: int = 1):
        self.field = super().compute(x, y, u, v, f, t, edge_order, -1)
        return self.field

    def compute_device(self, x, y, u=None, v=None, f=None, t=None, edge_order: int = 1):
        """Device-resident variant, see ``AttractionRate.compute_device``."""
        return self._eigen_field_device(x, y, u, v, f, t, edge_order, -1)


class _TrajectoryRepulsionRateRatio(EulerianDiagnostic2D):
    """rho-dot and nu-dot: quadratic forms of S along the velocity (R:eulerian.py:124-196);
    masked where the velocity vanishes (:188-192)."""

    def compute(self, x, y, u=None, v=None, f=None, t=None, edge_order: int = 1,
                rate_or_ratio: str = 'rate'):
        ud, vd = super().compute(x, y, u, v, f, t, edge_order)
        if rate_or_ratio == 'rate':
            mode = nat.EULER_RATE
        elif rate_or_ratio == 'ratio':
            mode = nat.EULER_RATIO
        else:
            raise ValueError('rate_or_ratio parameter must be either "rate" or "ratio".')
        field, mask, _ = self._euler(mode, ud, vd, edge_order)
        eng = Engine.get()
        self.field_device, self.mask_device = field, mask
        return np.ma.MaskedArray(eng.to_host(field), mask=eng.to_host(mask).astype(bool))


class TrajectoryRepulsionRate(_TrajectoryRepulsionRateRatio):
    """Trajectory repulsion rate (R:eulerian.py:199-226)."""

    def compute(self, x, y, u=None, v=None, f=None, t=None, edge_order: int = 1):
        self.field = super().compute(x, y, u, v, f, t, edge_order, 'rate')
        return self.field


class TrajectoryRepulsionRatio(_TrajectoryRepulsionRateRatio):
    """Trajectory repulsion ratio (R:eulerian.py:229-256)."""

    def compute(self, x, y, u=None, v=None, f=None, t=None, edge_order: int = 1):
        self.field = super().compute(x, y, u, v, f, t, edge_order, 'ratio')
        return self.field


class iLES(EulerianDiagnostic2D, RidgeExtractor2D):
    """Infinitesimal-time LCS: ridges of -s_1 (attracting) or s_n (repelling)
    (R:eulerian.py:259-351).  The default ``kind`` string is the reference's own
    (``'attacting'``, which its validation rejects); the sign flip compares ``kind`` without
    ``.lower()`` exactly as the reference does (:339)."""

    def compute(self, x, y, u=None, v=None, f=None, t=None, kind: str = 'attacting',
                edge_order: int = 1, percentile: float = None, force_eigenvectors: bool = True,
                debug: bool = False):
        if kind.lower() == 'attracting':
            mode = nat.EULER_S_MIN
        elif kind.lower() == 'repelling':
            mode = nat.EULER_S_MAX
        else:
            raise ValueError(
                f'kind: {kind}, unrecognized, please use either "attracting" or "repelling"')
        ud, vd = super().compute(x, y, u, v, f, t, edge_order)
        xi_mode = nat.XI_FORCED if force_eigenvectors else nat.XI_LAPACK
        field, _, xi = self._euler(mode, ud, vd, edge_order, negate=(kind == 'attracting'),
                                   xi_mode=xi_mode)
        eng = Engine.get()
        self.field = np.ma.MaskedArray(eng.to_host(field))
        self.iles = self.extract_ridges(field, xi, percentile, edge_order, debug)
        if debug:
            self.Xi_max = np.ma.MaskedArray(eng.to_host(xi))
        return self.iles
