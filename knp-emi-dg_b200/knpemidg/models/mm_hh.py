"""Hodgkin-Huxley membrane with Na/K pump and synaptic stimulus, SI units
(V, s, S/m^2).  Same equations and tables as the reference's
examples/idealized-geometries/mm_hh.py:7-161 (state/parameter order kept so
user code addressing columns by index keeps working)."""
import math
from knpemidg.models._protocol import build

STATES = [("m", 0.016648440745822956), ("h", 0.8542015627820805),
          ("n", 0.1882020248041632), ("V", -0.07438609374462003)]

PARAMETERS = [("g_Na_bar", 1200.0), ("g_K_bar", 360.0), ("g_leak_Na", 1.0), ("g_leak_K", 4.0),
              ("E_Na", 0.0), ("E_K", 0.0), ("Cm", 0.0), ("stim_amplitude", 0.0),
              ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0), ("K_e", 0.0), ("Na_i", 0.0),
              ("m_K", 2.0), ("m_Na", 7.7), ("I_max", 0.449), ("E_Cl", 0.0)]


def rhs(t, states, values, parameters):
    m = states[0]
    h = states[1]
    n = states[2]
    v = 1.0e3 * (states[3] + 65.0e-3)          # mV above rest (1952 convention)
    alpha_m = 0.1e3 * (25.0 - v) / (math.exp((25.0 - v) / 10.0) - 1)
    beta_m = 4.0e3 * math.exp(-v / 18.0)
    alpha_h = 0.07e3 * math.exp(-v / 20.0)
    beta_h = 1.0e3 / (math.exp((30.0 - v) / 10.0) + 1)
    alpha_n = 0.01e3 * (10.0 - v) / (math.exp((10.0 - v) / 10.0) - 1.0)
    beta_n = 0.125e3 * math.exp(-v / 80.0)
    values[0] = (1 - m) * alpha_m - m * beta_m
    values[1] = (1 - h) * alpha_h - h * beta_h
    values[2] = (1 - n) * alpha_n - n * beta_n
    g_syn = parameters[7] * math.exp(-math.fmod(t, 0.03) / 0.002) * (t < 125e-3)
    i_pump = parameters[15] / ((1 + parameters[13] / parameters[11]) ** 2
                               * (1 + parameters[14] / parameters[12]) ** 3)
    i_Na = (parameters[2] + parameters[0] * h * m ** 3 + g_syn) * (states[3] - parameters[4]) + 3 * i_pump
    i_K = (parameters[3] + parameters[1] * n ** 4) * (states[3] - parameters[5]) - 2 * i_pump
    parameters[8] = i_Na
    parameters[9] = i_K
    parameters[10] = 0.0
    values[3] = (-i_K - i_Na) / parameters[6]


globals().update(build(__name__, STATES, PARAMETERS, rhs))
