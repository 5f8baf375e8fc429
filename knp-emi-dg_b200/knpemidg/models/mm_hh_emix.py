"""Hodgkin-Huxley neuron membrane in mV / ms / mS cm^-2 units with pump and
periodic synaptic stimulus.  Same equations and tables as
examples/emix-simulations/mm_hh.py:7-162."""
import math
from knpemidg.models._protocol import build

STATES = [("m", 0.016651023270342777), ("h", 0.8541791472445746),
          ("n", 0.18821645700362638), ("V", -74.3848784437955)]

PARAMETERS = [("g_Na_bar", 120.0), ("g_K_bar", 36.0), ("g_leak_Na", 0.1), ("g_leak_K", 0.4),
              ("E_Na", 0.0), ("E_K", 0.0), ("Cm", 0.0), ("stim_amplitude", 0.0),
              ("I_ch_Na", 0.0), ("I_ch_K", 0.0), ("I_ch_Cl", 0.0), ("K_e", 0.0), ("Na_i", 0.0),
              ("m_K", 2.0), ("m_Na", 7.7), ("I_max", 44.9), ("E_Cl", 0.0)]


def rhs(t, states, values, parameters):
    m = states[0]
    h = states[1]
    n = states[2]
    V = states[3]
    alpha_m = 0.1 * (V + 40.0) / (1.0 - math.exp(-(V + 40.0) / 10.0))
    beta_m = 4.0 * math.exp(-(V + 65.0) / 18.0)
    alpha_h = 0.07 * math.exp(-(V + 65.0) / 20.0)
    beta_h = 1.0 / (1.0 + math.exp(-(V + 35.0) / 10.0))
    alpha_n = 0.01 * (V + 55.0) / (1.0 - math.exp(-(V + 55.0) / 10.0))
    beta_n = 0.125 * math.exp(-(V + 65) / 80.0)
    values[0] = (1 - m) * alpha_m - m * beta_m
    values[1] = (1 - h) * alpha_h - h * beta_h
    values[2] = (1 - n) * alpha_n - n * beta_n
    g_syn = parameters[7] * math.exp(-math.fmod(t, 20.0) / 2.0)
    i_pump = parameters[15] / ((1 + parameters[13] / parameters[11]) ** 2
                               * (1 + parameters[14] / parameters[12]) ** 3)
    i_Na = (parameters[2] + parameters[0] * h * m ** 3 + g_syn) * (V - parameters[4]) + 3 * i_pump
    i_K = (parameters[3] + parameters[1] * n ** 4) * (V - parameters[5]) - 2 * i_pump
    parameters[8] = i_Na
    parameters[9] = i_K
    parameters[10] = 0.0
    values[3] = (-i_K - i_Na) / parameters[6]


globals().update(build(__name__, STATES, PARAMETERS, rhs))
