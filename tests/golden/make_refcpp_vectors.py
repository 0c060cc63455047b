#!/usr/bin/env python
"""Regenerates tests/golden/refcpp_vectors.json from the reference's OWN native benchmark program
(benchmarks/cpp of ForwardDiff.jl, built from /root/reference into oracle/_ref/libfdref.so by `make -C oracle ref`).

    python tests/golden/make_refcpp_vectors.py

The vectors are outputs of the reference's code, not of this repository's: gradient<F<DualK>>(result, dualvec, input)
for F in {rosenbrock, ackley}, K = 1..5, on a seeded input of 60 elements (60 % K == 0 for every K, as the reference's
chunk loop requires, benchmarks.h:120-227).  Doubles are stored as hex strings so the fixture is bit-exact.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc


def main():
    rb = orc.RefBench()
    assert rb.testgrad() == 0                              # the reference's own known-answer asserts pass
    x = np.random.default_rng(2024).uniform(0.25, 1.75, 60)
    out = {"_comment": "outputs of /root/reference/benchmarks/cpp (g++ -std=c++11 -O2), see make_refcpp_vectors.py",
           "x": [float(v).hex() for v in x], "cases": []}
    for fn in ("rosenbrock", "ackley"):
        out["value_" + fn] = float(rb.value(fn, x)).hex()
        for K in (1, 2, 3, 4, 5):
            g = rb.gradient(fn, x, K)
            out["cases"].append({"fn": fn, "chunk": K, "gradient": [float(v).hex() for v in g]})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refcpp_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
