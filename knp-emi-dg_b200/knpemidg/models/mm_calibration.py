"""Three-compartment calibration system (mV / ms units): one Hodgkin-Huxley neuron membrane and
one glial membrane (Na leak, Kir4.1, Na/K pumps on both) exchanging K+ and Na+ with a shared
extracellular volume - 11 states, integrated to its steady state to obtain consistent initial
data for the EMIx runs.  Same equations and tables as
examples/emix-simulations/mm_calibration.py:4-255 (driver: run_calibration.py:13-64, whose
outcome is hard-coded as the initial values of examples/emix-simulations/mm_hh.py:11-14).
No PDE coupling: there is no state called V and no I_ch_* parameter."""
import math
from knpemidg.models._protocol import build

STATES = [("m", 0.01), ("h", 0.85), ("n", 0.18), ("V_n", -74.38), ("V_g", -83.08),
          ("K_e", 3.32), ("K_n", 124.15), ("K_g", 102.75),
          ("Na_e", 100.71), ("Na_n", 12.83), ("Na_g", 12.39)]

PARAMETERS = [("g_Na_bar", 120.0), ("g_K_bar", 36.0), ("g_leak_Na_n", 0.1), ("g_leak_K_n", 0.4),
              ("g_leak_Na_g", 0.1), ("g_leak_K_g", 1.7), ("Cm", 2.0), ("stim_amplitude", 0.0),
              ("m_K", 2.0), ("m_Na", 7.7), ("I_max_n", 44.9), ("I_max_g", 50.0)]


def rhs(t, states, values, parameters):
    temperature = 300e3
    R = 8.314e3
    F = 96485e3
    RTF = R * temperature / F
    # compartment geometry: half of the intracellular volume each for neuron and glia
    vol_i = 3.42e-11 / 2.0
    vol_e = 7.08e-11
    area = 2.29e-6
    K_g_ref = 102.74050220804774
    K_e_ref = 3.32597273958481

    gate_m = states[0]
    gate_h = states[1]
    gate_n = states[2]
    V_n = states[3]
    V_g = states[4]
    K_e = states[5]
    K_n = states[6]
    K_g = states[7]
    Na_e = states[8]
    Na_n = states[9]
    Na_g = states[10]

    # Nernst potentials of both membranes
    E_Na_n = RTF * math.log(Na_e / Na_n)
    E_K_n = RTF * math.log(K_e / K_n)
    E_Na_g = RTF * math.log(Na_e / Na_g)
    E_K_g = RTF * math.log(K_e / K_g)
    E_K_ref = RTF * math.log(K_e_ref / K_g_ref)

    # Hodgkin-Huxley gates of the neuron
    a_m = 0.1 * (V_n + 40.0) / (1.0 - math.exp(-(V_n + 40.0) / 10.0))
    b_m = 4.0 * math.exp(-(V_n + 65.0) / 18.0)
    a_h = 0.07 * math.exp(-(V_n + 65.0) / 20.0)
    b_h = 1.0 / (1.0 + math.exp(-(V_n + 35.0) / 10.0))
    a_n = 0.01 * (V_n + 55.0) / (1.0 - math.exp(-(V_n + 55.0) / 10.0))
    b_n = 0.125 * math.exp(-(V_n + 65) / 80.0)
    values[0] = (1 - gate_m) * a_m - gate_m * b_m
    values[1] = (1 - gate_h) * a_h - gate_h * b_h
    values[2] = (1 - gate_n) * a_n - gate_n * b_n

    stim = parameters[7] * math.exp(-(t % 20.0) / 2.0)
    pump_n = parameters[10] / ((1 + parameters[8] / K_e) ** 2 * (1 + parameters[9] / Na_n) ** 3)
    pump_g = parameters[11] / ((1 + parameters[8] / K_e) ** 2 * (1 + parameters[9] / Na_g) ** 3)

    # Kir4.1 conductance factor
    A = 1 + math.exp(18.4 / 42.4)
    B = 1 + math.exp(-(0.1186e3 + E_K_ref) / 0.0441e3)
    C = 1 + math.exp((V_g - E_K_g + 0.0185e3) / 0.0425e3)
    D = 1 + math.exp(-(0.1186e3 + V_g) / 0.0441e3)
    g_Kir = math.sqrt(K_e / K_e_ref) * (A * B) / (C * D)
    i_Kir = parameters[5] * g_Kir * (V_g - E_K_g)

    i_Na_n = (parameters[2] + parameters[0] * gate_h * gate_m ** 3 + stim) * (V_n - E_Na_n) + 3 * pump_n
    i_K_n = (parameters[3] + parameters[1] * gate_n ** 4) * (V_n - E_K_n) - 2 * pump_n
    i_Na_g = parameters[4] * (V_g - E_Na_g) + 3 * pump_g
    i_K_g = i_Kir - 2 * pump_g

    values[3] = (-i_K_n - i_Na_n) / parameters[6]
    values[4] = (-i_K_g - i_Na_g) / parameters[6]
    # ion budgets of the three compartments
    values[5] = i_K_n * area / (F * vol_e) + i_K_g * area / (F * vol_e)
    values[6] = -i_K_n * area / (F * vol_i)
    values[7] = -i_K_g * area / (F * vol_i)
    values[8] = i_Na_n * area / (F * vol_e) + i_Na_g * area / (F * vol_e)
    values[9] = -i_Na_n * area / (F * vol_i)
    values[10] = -i_Na_g * area / (F * vol_i)


globals().update(build(__name__, STATES, PARAMETERS, rhs))
