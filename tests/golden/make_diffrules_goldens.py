#!/usr/bin/env python
"""Independent goldens for the DiffRules derivative formulas.

The oracle (oracle/oracle.cpp) and the device header (csrc/fdb_dual.cuh) restate the same DiffRules 1.x expressions, so a
shared misreading of a rule would pass every kernel-vs-oracle test.  This script pins each rule to mathematics instead:
for every unary op of `fdb_op1` and every binary op of `fdb_op2` (Dual (x) Dual, Dual (x) Real, Real (x) Dual) it evaluates
the function and its exact derivative(s) with mpmath at 50 significant digits on well-conditioned inputs and stores the
correctly rounded Float64 results as hex strings in tests/golden/diffrules_goldens.json.  The reference's own check of
these rules is test/DualTest.jl:540-613 (value == f(x), partial == the DiffRules expression) -- relative to DiffRules
itself; these goldens are relative to the true derivative.

    python tests/golden/make_diffrules_goldens.py          (needs mpmath; deterministic)

Tests (tests/test_diffrules_goldens.py): value within 2 ulp of the true value, partial within 4 ulp of
p x (true derivative) -- for the CPU oracle in the `not gpu` suite and for the device kernels in the `gpu` suite.
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))


def hx(v):
    return float(v).hex()


def grid(lo, hi, k, seed):
    rng = np.random.default_rng(seed)
    return [float(v) for v in rng.uniform(lo, hi, k)]


LN2, LN10 = mp.log(2), mp.log(10)
# name -> (inputs, f, f')
UNARY = {
    "neg": (grid(-3, 3, 12, 1), lambda x: -x, lambda x: mp.mpf(-1)),
    "exp": (grid(-3, 3, 12, 2), mp.exp, mp.exp),
    "log": (grid(0.2, 5, 12, 3), mp.log, lambda x: 1 / x),
    "sin": (grid(-3, 3, 12, 4), mp.sin, mp.cos),
    "cos": (grid(-3, 3, 12, 5), mp.cos, lambda x: -mp.sin(x)),
    "tanh": (grid(-1, 1, 12, 6), mp.tanh, lambda x: 1 - mp.tanh(x) ** 2),
    "sqrt": (grid(0.1, 9, 12, 7), mp.sqrt, lambda x: 1 / (2 * mp.sqrt(x))),
    "abs": (grid(-3, 3, 12, 8), abs, lambda x: mp.sign(x)),
    "abs2": (grid(-3, 3, 12, 9), lambda x: x * x, lambda x: 2 * x),
    "inv": (grid(0.2, 4, 6, 10) + grid(-4, -0.2, 6, 11), lambda x: 1 / x, lambda x: -1 / (x * x)),
    "pow0": (grid(-3, 3, 6, 12), lambda x: mp.mpf(1), lambda x: mp.mpf(0)),
    "pow1": (grid(-3, 3, 6, 13), lambda x: x, lambda x: mp.mpf(1)),
    "pow2": (grid(-3, 3, 12, 14), lambda x: x ** 2, lambda x: 2 * x),
    "pow3": (grid(-3, 3, 12, 15), lambda x: x ** 3, lambda x: 3 * x ** 2),
    "tan": (grid(-1.2, 1.2, 12, 16), mp.tan, lambda x: 1 + mp.tan(x) ** 2),
    "exp2": (grid(-3, 3, 12, 17), lambda x: mp.mpf(2) ** x, lambda x: mp.mpf(2) ** x * LN2),
    "log2": (grid(0.2, 5, 12, 18), lambda x: mp.log(x) / LN2, lambda x: 1 / (x * LN2)),
    "log10": (grid(0.2, 5, 12, 19), mp.log10, lambda x: 1 / (x * LN10)),
    "log1p": (grid(-0.5, 3, 12, 20), mp.log1p, lambda x: 1 / (1 + x)),
    "expm1": (grid(-2, 2, 12, 21), mp.expm1, mp.exp),
    "sinh": (grid(-3, 3, 12, 22), mp.sinh, mp.cosh),
    "cosh": (grid(-3, 3, 12, 23), mp.cosh, mp.sinh),
    "asin": (grid(-0.7, 0.7, 12, 24), mp.asin, lambda x: 1 / mp.sqrt(1 - x * x)),
    "acos": (grid(-0.7, 0.7, 12, 25), mp.acos, lambda x: -1 / mp.sqrt(1 - x * x)),
    "atan": (grid(-3, 3, 12, 26), mp.atan, lambda x: 1 / (1 + x * x)),
    "cbrt": (grid(0.1, 9, 12, 27), mp.cbrt, lambda x: 1 / (3 * mp.cbrt(x) ** 2)),
}


def pairs(lox, hix, loy, hiy, k, seed):
    rng = np.random.default_rng(seed)
    return [(float(a), float(b)) for a, b in zip(rng.uniform(lox, hix, k), rng.uniform(loy, hiy, k))]


# name -> (input pairs (a, b), f(a, b), df/da, df/db)
BINARY = {
    "add": (pairs(-3, 3, -3, 3, 10, 31), lambda a, b: a + b, lambda a, b: mp.mpf(1), lambda a, b: mp.mpf(1)),
    "sub": (pairs(-3, 3, -3, 3, 10, 32), lambda a, b: a - b, lambda a, b: mp.mpf(1), lambda a, b: mp.mpf(-1)),
    "mul": (pairs(-3, 3, -3, 3, 10, 33), lambda a, b: a * b, lambda a, b: b, lambda a, b: a),
    "div": (pairs(-3, 3, 0.3, 3, 10, 34), lambda a, b: a / b, lambda a, b: 1 / b, lambda a, b: -a / (b * b)),
    "pow": (pairs(0.5, 2.5, -2, 2, 10, 35), lambda a, b: a ** b, lambda a, b: b * a ** (b - 1), lambda a, b: a ** b * mp.log(a)),
    "max": (pairs(-3, 3, -3, 3, 10, 36), lambda a, b: max(a, b), lambda a, b: mp.mpf(1 if a > b else 0), lambda a, b: mp.mpf(0 if a > b else 1)),
    "min": (pairs(-3, 3, -3, 3, 10, 37), lambda a, b: min(a, b), lambda a, b: mp.mpf(1 if a < b else 0), lambda a, b: mp.mpf(0 if a < b else 1)),
    # atan(y, x): first argument is y
    "atan2": (pairs(-3, 3, 0.2, 3, 10, 38), lambda y, x: mp.atan2(y, x), lambda y, x: x / (x * x + y * y), lambda y, x: -y / (x * x + y * y)),
    "hypot": (pairs(0.2, 3, 0.2, 3, 10, 39), lambda a, b: mp.hypot(a, b), lambda a, b: a / mp.hypot(a, b), lambda a, b: b / mp.hypot(a, b)),
}


def main():
    out = {"generator": "tests/golden/make_diffrules_goldens.py", "mpmath": mp.__version__, "dps": mp.mp.dps,
           "format": "hex Float64; value/derivatives are the 50-digit results rounded to nearest", "unary": {}, "binary": {}}
    for name, (xs, f, df) in UNARY.items():
        out["unary"][name] = [{"x": hx(x), "value": hx(f(mp.mpf(x))), "deriv": hx(df(mp.mpf(x)))} for x in xs]
    for name, (ps, f, dfa, dfb) in BINARY.items():
        out["binary"][name] = [{"a": hx(a), "b": hx(b), "value": hx(f(mp.mpf(a), mp.mpf(b))), "da": hx(dfa(mp.mpf(a), mp.mpf(b))),
                                "db": hx(dfb(mp.mpf(a), mp.mpf(b)))} for a, b in ps]
    path = os.path.join(HERE, "diffrules_goldens.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    n = sum(len(v) for v in out["unary"].values()) + sum(len(v) for v in out["binary"].values())
    print(f"wrote {n} golden points to {path}")


if __name__ == "__main__":
    main()
