"""Decay widths (MeV).  Scalar, evaluated once per (m_a, g) on the host exactly as the reference does
(decay.py:10-52); the per-sample lifetime / survival arithmetic that uses them runs in the kernels."""
import numpy as np
from numpy import arcsin, log, pi, power, sqrt

from .constants import ALPHA, M_E


def W_gg(g_agamma, ma):
    """a -> gamma gamma, g in MeV^-1 (decay.py:10-13)."""
    return g_agamma**2 * ma**3 / (64*pi)


def W_ff(g, mf, ma):
    """a -> f fbar (decay.py:18-21)."""
    return g**2 * ma * sqrt(1 - (2 * mf / ma)**2) / (8 * pi) \
        if 1 - 4 * (mf / ma) ** 2 > 0 else 0.0


def W_ee(g_ae, ma):
    """a -> e+ e- (decay.py:26-29); zero below threshold."""
    return g_ae**2 * ma * sqrt(1 - (2 * M_E / ma)**2) / (8 * pi) \
        if 1 - 4 * (M_E / ma) ** 2 > 0 else 0.0


def fp2(tau):
    """Loop function (decay.py:35-37)."""
    return arcsin(1/sqrt(tau))**2 if tau >= 1 \
        else pi**2 / 4 + log((1+sqrt(1-tau))/(1-sqrt(1-tau)))**2 / 4


def b1(tau):
    return 1 - tau*fp2(tau)


def W_gg_loop(g_af, ma, mf):
    """a -> gamma gamma through a fermion loop (decay.py:48-52)."""
    g_agamma_eff = g_af * ALPHA * b1(power(2*mf/ma, 2)) / (4*pi)
    return W_gg(g_agamma_eff, ma)


_GLUDOM = None


def W_agg_hadronic(f_a, ma):
    """a -> hadrons total width from the tabulated gluon-dominance widths (decay.py:65-72); f_a, ma in MeV."""
    global _GLUDOM
    if _GLUDOM is None:
        from .materials import data_path
        _GLUDOM = np.genfromtxt(data_path("gluon_dominance_widths_nogamma.txt"))
    eV_fact = 1e-9
    Lambda_scale = (32 * pi**2 * f_a / 1e6)**2  # data is normalized to an EFT scale of 1 TeV
    return np.interp(ma*1e-3, _GLUDOM[:, 0], _GLUDOM[:, 1]) * eV_fact / Lambda_scale
