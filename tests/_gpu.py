"""Helpers for the -m gpu parity tests: build an Engine (through the C ABI) from the engine inputs of a case."""
import numpy as np


def make_engine(inp, device=0):
    import hpv_b200
    eng = hpv_b200.Engine(device)
    eng.set_network(inp["layers"], inp["act"])
    eng.set_quadrature(inp["xi"], inp["w"])
    eng.set_test_tables(inp["T"], inp["D1"], inp["D2"], inp["d1b"])
    eng.set_form(inp["problem"], inp["var_form"], inp.get("V", 1.0))
    eng.set_elements(inp["lo"], inp["hi"], inp["ntx"], inp["nty"], inp["F"], inp.get("ntest"))
    eng.set_params(inp["theta"], inp.get("eps", 0.0))
    return eng


# gradient tolerance (relative to max |grad|): fp32 conditioning of the form, see DESIGN.md "Precision"
GRAD_RTOL = {"p1d_vf2": 5e-4, "p1d_vf3": 3e-2}
