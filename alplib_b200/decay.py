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
