from .lagrangian import FTLE, LCS  # noqa: F401
from .eulerian import (  # noqa: F401
    AttractionRate, RepulsionRate, iLES, TrajectoryRepulsionRate, TrajectoryRepulsionRatio
)
