#!/usr/bin/env python3
"""Extract the reference's golden vectors into small committed fixtures.

Run in the build container (where /root/reference exists); the outputs travel
with the repo so no test ever reads /root/reference at run time.

  tests/golden/pulse_regression.npz
      RHO_t0/100/200 (5000,), VEL_t0/100/200 (15000,) float64, exactly as printed
      in /root/reference/tests/pulseRegression.h:6-22 (6 significant digits;
      order k-outer, j, i-inner, n innermost -- tests/catch2RegressionTests.cpp:44-52)
  tests/golden/mode_matrices.npz
      MODE_MATRIX, MODE_MATRIX_INVERSE (15,15) evaluated from the literals at
      /root/reference/src/AmrSim.cpp:1037-1073; DELTA (:1033-1035);
      CX, CY, CZ, W from /root/reference/include/d3q15_bgk.h:14-28.
"""
import os
import re
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def braces(text, start):
    i = text.index("{", start)
    depth, j = 0, i
    while True:
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                return text[i:j + 1]
        j += 1


def to_list(src):
    return eval(src.replace("{", "[").replace("}", "]"), {"__builtins__": {}}, {})


def main():
    txt = open(os.path.join(REF, "tests/pulseRegression.h")).read()
    out = {}
    for name in ("RHO_t0", "VEL_t0", "RHO_t100", "VEL_t100", "RHO_t200", "VEL_t200"):
        m = re.search(r"\b%s\s*=" % name, txt)
        out[name] = np.array(to_list(braces(txt, m.end())), dtype=np.float64)
    assert all(out[k].size == 5000 for k in out if k.startswith("RHO"))
    assert all(out[k].size == 15000 for k in out if k.startswith("VEL"))
    np.savez_compressed(os.path.join(HERE, "pulse_regression.npz"), **out)

    src = open(os.path.join(REF, "src/AmrSim.cpp")).read()
    mm = {}
    for name in ("MODE_MATRIX", "MODE_MATRIX_INVERSE"):
        m = re.search(r"AmrSim::%s\[NMODES\]\[NMODES\]\s*=" % name, src)
        mm[name] = np.array(to_list(braces(src, m.end())), dtype=np.float64)
        assert mm[name].shape == (15, 15)
    m = re.search(r"AmrSim::DELTA\[NDIMS\]\[NDIMS\]\s*=", src)
    mm["DELTA"] = np.array(to_list(braces(src, m.end()).replace("NMODES", "15")), dtype=np.float64)
    hdr = open(os.path.join(REF, "include/d3q15_bgk.h")).read()
    for name in ("CX", "CY", "CZ", "W"):
        m = re.search(r"\b%s\s*=" % name, hdr)
        mm[name] = np.array(to_list(braces(hdr, m.end())), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "mode_matrices.npz"), **mm)
    print({k: v.shape for k, v in {**out, **mm}.items()})


if __name__ == "__main__":
    main()
