"""Astrocyte membrane: Na, Cl leak, Kir4.1, saturating Na/K pump (mV / ms).
Same equations and tables as
examples/local-astrocyte-depolarization/mm_glial.py:6-190."""
import math
from knpemidg.models._protocol import build

STATES = [("V", -85.85765274084892)]

PARAMETERS = [("g_leak_Cl", 0.05), ("g_leak_Na", 0.1), ("g_leak_K", 1.696),
              ("E_Cl", 0.0), ("E_Na", 0.0), ("E_K", 0.0), ("Cm", 0.0), ("stim_amplitude", 0.0),
              ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0), ("K_e", 0.0), ("Na_i", 0.0),
              ("m_K", 1.5), ("m_Na", 10.0), ("I_max", 10.75975),
              ("K_e_init", 3.092970607490389), ("K_i_init", 99.3100014897692)]


def rhs(t, states, values, parameters):
    V = states[0]
    E_K = parameters[5]
    K_e = parameters[11]
    Na_i = parameters[12]
    temperature = 307e3
    R = 8.315e3
    F = 96500e3
    i_pump = parameters[15] * (K_e / (K_e + parameters[13])) \
        * (Na_i ** 1.5 / (Na_i ** 1.5 + parameters[14] ** 1.5))
    E_K_init = R * temperature / F * math.log(parameters[16] / parameters[17])
    dphi = V - E_K
    A = 1 + math.exp(18.4 / 42.4)
    B = 1 + math.exp(-(0.1186e3 + E_K_init) / 0.0441e3)
    C = 1 + math.exp((dphi + 0.0185e3) / 0.0425e3)
    D = 1 + math.exp(-(0.1186e3 + V) / 0.0441e3)
    g_Kir = math.sqrt(K_e / parameters[16]) * (A * B) / (C * D)
    i_Kir = parameters[2] * g_Kir * (V - E_K)
    i_Na = parameters[1] * (V - parameters[4]) + 3 * i_pump
    i_K = i_Kir - 2 * i_pump
    i_Cl = parameters[0] * (V - parameters[3])
    parameters[8] = i_Na
    parameters[9] = i_K
    parameters[10] = i_Cl
    values[0] = (-i_K - i_Na - i_Cl) / parameters[6]


globals().update(build(__name__, STATES, PARAMETERS, rhs))
