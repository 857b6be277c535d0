"""The radiation problem the reference's test_driver sets up on its tiled meshes
(driver/test_driver.cc:1163-1287, 1829-1842; interface/TetonConduitInterface.cc:1479-1484):
log-spaced group bounds on [1e-6, 1e2], Tr0 = 0.05, Te0 = 0.5, rho = 1.31, cv = 0.501,
all-vacuum boundaries, fixed dt = 1e-3.  In the mini-app build sigA = sigS = 0 and
STotal = 0, so Sigt = tau = 1/(c dt) (control/setTotalOpacity.F90:52)."""
from __future__ import annotations

import math

import numpy as np

SPEED_LIGHT = 299.792458             # mods/radconstant_mod.F90:27
RAD_CONSTANT = 0.013720169264801055  # mods/radconstant_mod.F90:28
TR0, TE0, RHO, CV = 0.05, 0.5, 1.31, 0.501
DT = 1.0e-3
TFLOOR = 1.0e-5


def group_bounds(ngr: int) -> np.ndarray:
    """test_driver.cc:1173-1183: exp(ln 1e-6 + g ln(1e8)/G)."""
    lo, hi = math.log(1.0e-6), math.log(1.0e2)
    return np.exp(lo + np.arange(ngr + 1) * (hi - lo) / ngr)


def wtiso(ndim: int) -> float:
    return 1.0 / (4.0 * math.pi) if ndim == 3 else 1.0 / (2.0 * math.pi)   # Size_mod.F90:278-281


def tau(dt: float = DT) -> float:
    return 1.0 / (SPEED_LIGHT * dt)   # initializeSets.F90:81
