"""Hodgkin-Huxley membrane with Na/K pump, no stimulus, SI units.  Same
equations and tables as examples/idealized-geometries/mm_hh_no_stim.py:7-159
(identical copy in examples/rat-neuron/mm_hh_no_stim.py)."""
import math
from knpemidg.models._protocol import build
from knpemidg.models.mm_hh import STATES, PARAMETERS


def rhs(t, states, values, parameters):
    m = states[0]
    h = states[1]
    n = states[2]
    v = 1.0e3 * (states[3] + 65.0e-3)
    alpha_m = 0.1e3 * (25.0 - v) / (math.exp((25.0 - v) / 10.0) - 1)
    beta_m = 4.0e3 * math.exp(-v / 18.0)
    alpha_h = 0.07e3 * math.exp(-v / 20.0)
    beta_h = 1.0e3 / (math.exp((30.0 - v) / 10.0) + 1)
    alpha_n = 0.01e3 * (10.0 - v) / (math.exp((10.0 - v) / 10.0) - 1.0)
    beta_n = 0.125e3 * math.exp(-v / 80.0)
    values[0] = (1 - m) * alpha_m - m * beta_m
    values[1] = (1 - h) * alpha_h - h * beta_h
    values[2] = (1 - n) * alpha_n - n * beta_n
    i_pump = parameters[15] / ((1 + parameters[13] / parameters[11]) ** 2
                               * (1 + parameters[14] / parameters[12]) ** 3)
    i_Na = (parameters[2] + parameters[0] * h * m ** 3) * (states[3] - parameters[4]) + 3 * i_pump
    i_K = (parameters[3] + parameters[1] * n ** 4) * (states[3] - parameters[5]) - 2 * i_pump
    parameters[8] = i_Na
    parameters[9] = i_K
    parameters[10] = 0.0
    values[3] = (-i_K - i_Na) / parameters[6]


globals().update(build(__name__, STATES, PARAMETERS, rhs))
