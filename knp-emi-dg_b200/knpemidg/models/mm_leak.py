"""Passive leak membrane with Na/K pump and synaptic stimulus, SI units.
Same equations and tables as examples/rat-neuron/mm_leak.py:7-133."""
import math
from knpemidg.models._protocol import build

STATES = [("V", -0.07438609374462003)]

PARAMETERS = [("g_leak_Na", 1.0), ("g_leak_K", 4.0), ("E_Na", 0.0), ("E_K", 0.0), ("Cm", 0.0),
              ("stim_amplitude", 0.0), ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0),
              ("K_e", 0.0), ("Na_i", 0.0), ("m_K", 2.0), ("m_Na", 7.7), ("I_max", 0.449),
              ("E_Cl", 0.0)]


def rhs(t, states, values, parameters):
    g_syn = parameters[5] * math.exp(-math.fmod(t, 0.03) / 0.002)
    i_pump = parameters[13] / ((1 + parameters[11] / parameters[9]) ** 2
                               * (1 + parameters[12] / parameters[10]) ** 3)
    i_Na = (parameters[0] + g_syn) * (states[0] - parameters[2]) + 3 * i_pump
    i_K = parameters[1] * (states[0] - parameters[3]) - 2 * i_pump
    parameters[6] = i_Na
    parameters[7] = i_K
    parameters[8] = 0.0
    values[0] = (-i_K - i_Na) / parameters[4]


globals().update(build(__name__, STATES, PARAMETERS, rhs))
