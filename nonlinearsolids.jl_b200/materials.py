"""Host-side mirror of src/materials/: parameter records selecting a device constitutive law.

The constitutive arithmetic itself runs inside the element kernels (csrc/kernels_assembly.cu);
the scalar `sigma`/`dsigma` methods here restate the closed forms for host-side post-processing.
"""
import math

from . import _lib


class AbstractMaterial:
    kind = None
    rho = 1.0

    def density(self):
        return self.rho

    def params(self):
        raise NotImplementedError("Not implemented")

    has_state = False
    is_velocitydependent = False


class LinearElasticity1D(AbstractMaterial):
    """sigma = E*eps (linearelasticity.jl:11-19)."""
    kind = _lib.MAT_LINEAR_ELASTIC_1D

    def __init__(self, E, rho=1.0):
        self.E, self.rho = float(E), float(rho)

    def params(self):
        return [self.E]

    def sigma(self, eps):
        return self.E * eps

    def dsigma(self, eps):
        return self.E


class ExponentialMaterial(AbstractMaterial):
    """sigma = sigma_sat*(1 - exp(-b*eps)) (exponential.jl:11-21)."""
    kind = _lib.MAT_EXPONENTIAL_1D

    def __init__(self, sigma_sat, b, rho=1.0):
        self.sigma_sat, self.b, self.rho = float(sigma_sat), float(b), float(rho)

    def params(self):
        return [self.sigma_sat, self.b]

    def sigma(self, eps):
        return self.sigma_sat * (1.0 - math.exp(-self.b * eps))

    def dsigma(self, eps):
        return self.b * self.sigma_sat * math.exp(-self.b * eps)


class LinearElastoPlasticity1D(AbstractMaterial):
    """1-D elastoplasticity, isotropic + kinematic hardening (linearelastoplasticity.jl:5-20)."""
    kind = _lib.MAT_ELASTOPLASTIC_1D
    has_state = True

    def __init__(self, E, H_kappa, H_alpha, kappa0, alpha0, Eep=None, rho=1.0):
        self.E, self.Hk, self.Ha, self.k0, self.a0 = map(float, (E, H_kappa, H_alpha, kappa0, alpha0))
        self.Eep = self.E * (self.Hk + self.Ha) / (self.E + self.Hk + self.Ha) if Eep is None else float(Eep)
        self.rho = float(rho)

    def params(self):
        return [self.E, self.Hk, self.Ha, self.k0, self.a0, self.Eep]

    def sigma(self, eps):
        raise RuntimeError("Use getstate(fem, 'sig') to get sigma")

    def dsigma(self, eps):
        raise RuntimeError("Use getstate(fem, 'dsig') to get dsigma")


class NeoHookean(AbstractMaterial):
    """Compressible Neo-Hookean, psi = mu/2 (I1-3) - mu lnJ + lam/2 (lnJ)^2 (T2 extension).

    2-D meshes use plane strain (F33 = 1)."""
    kind = _lib.MAT_NEOHOOKEAN

    def __init__(self, mu=None, lam=None, E=None, nu=None, rho=1.0):
        if mu is None:
            mu = E / (2.0 * (1.0 + nu))
            lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu))
        self.mu, self.lam, self.rho = float(mu), float(lam), float(rho)

    def params(self):
        return [self.mu, self.lam]


class J2Plasticity(AbstractMaterial):
    """Small-strain J2 elastoplasticity with linear isotropic hardening on Q1 quads (plane strain) and hexahedra: the
    stateful-material protocol of LinearElastoPlasticity1D (linearelastoplasticity.jl:60-117) generalised to 2-D / 3-D
    (SURVEY 8 f-2). State per quadrature point: plastic strain `epsp_xx ... epsp_xz` and `alpha` (trial; `_n` = committed)."""
    kind = _lib.MAT_J2_SMALL_STRAIN
    has_state = True

    def __init__(self, E, nu, sigma_y, H, rho=1.0):
        self.E, self.nu, self.sigma_y, self.H, self.rho = map(float, (E, nu, sigma_y, H, rho))
        self.K = self.E / (3.0 * (1.0 - 2.0 * self.nu))
        self.G = self.E / (2.0 * (1.0 + self.nu))

    def params(self):
        return [self.K, self.G, self.sigma_y, self.H]


def initialize_state(material, fem):
    """initialize_state!(material, fem) (linearelastoplasticity.jl:106-117)."""
    fem.handle(material).call("nls_init_state")


def save_state(material, fem, U, ctx=None):
    """save_state!(material, fem, U, ctx) (linearelastoplasticity.jl:96-103)."""
    if material.has_state:
        import numpy as np
        U = np.ascontiguousarray(U, dtype=np.float64)
        fem.handle(material).call("nls_commit_state", _lib.dptr(U))
