"""Glial membrane: Na leak, Kir4.1 inward-rectifying K channel, Na/K pump
(mV / ms units).  Same equations and tables as
examples/emix-simulations/mm_glial.py:6-170."""
import math
from knpemidg.models._protocol import build

STATES = [("V", -83.08511451850003)]

PARAMETERS = [("g_Na_bar", 0.0), ("g_K_bar", 0.0), ("g_leak_Na", 0.1), ("g_leak_K", 1.7),
              ("E_Na", 0.0), ("E_K", 0.0), ("Cm", 0.0), ("stim_amplitude", 0.0),
              ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0), ("K_e", 0.0), ("Na_i", 0.0),
              ("m_K", 2.0), ("m_Na", 7.7), ("I_max", 50.0), ("K_e_init", 3.32597273958481),
              ("K_i_init", 102.74050220804774), ("E_Cl", 0.0)]


def rhs(t, states, values, parameters):
    V = states[0]
    K_e = parameters[11]
    i_pump = parameters[15] / ((1 + parameters[13] / K_e) ** 2
                               * (1 + parameters[14] / parameters[12]) ** 3)
    temperature = 300e3
    R = 8.314e3
    F = 96485e3
    E_K_init = R * temperature / F * math.log(parameters[16] / parameters[17])
    dphi = V - parameters[5]
    A = 1 + math.exp(18.4 / 42.4)
    B = 1 + math.exp(-(0.1186e3 + E_K_init) / 0.0441e3)
    C = 1 + math.exp((dphi + 0.0185e3) / 0.0425e3)
    D = 1 + math.exp(-(0.1186e3 + V) / 0.0441e3)
    g_Kir = math.sqrt(K_e / parameters[16]) * (A * B) / (C * D)
    i_Kir = parameters[3] * g_Kir * (V - parameters[5])
    i_Na = parameters[2] * (V - parameters[4]) + 3 * i_pump
    i_K = i_Kir - 2 * i_pump
    parameters[8] = i_Na
    parameters[9] = i_K
    parameters[10] = 0.0
    values[0] = (-i_K - i_Na) / parameters[6]


globals().update(build(__name__, STATES, PARAMETERS, rhs))
