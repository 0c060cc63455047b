"""Target functions.

`rosenbrock`, `ackley`, ... are *catalogue* targets: the drivers recognise them
(attribute `fdb_catalogue`) and run the fused, batched device sweep, the way the
Julia shim specialises `chunk_mode_gradient(::typeof(rosenbrock), ::B200Array, cfg)`.

The `*_generic` functions are ordinary callables written against B200DualArray
arithmetic (broadcast + sum), i.e. what an arbitrary user `f` looks like; they
run through the reference's chunk loop on materialised device duals.
"""
import math

from . import array as A


class CatalogueFn:
    def __init__(self, name, doc, output_length=None, **params):
        self.fdb_catalogue = name
        self.__name__ = name
        self.__doc__ = doc
        self.params = params
        self._outlen = output_length

    def output_length(self, n):
        return self._outlen(n) if self._outlen else None

    def with_params(self, **params):
        """A new target object carrying Float64 parameters (A, b, alpha)."""
        return CatalogueFn(self.fdb_catalogue, self.__doc__, self._outlen, **params)

    def __call__(self, *args):
        raise TypeError(f"{self.__name__} is a device catalogue target; pass it to gradient/jacobian/hessian")

    def __repr__(self):
        return f"<fdb200 catalogue target {self.__name__}>"


rosenbrock = CatalogueFn("rosenbrock", "docs/src/user/advanced.md:36-44 (= DiffTests.rosenbrock_1)")
ackley = CatalogueFn("ackley", "DiffTests.ackley; benchmarks/cpp/benchmarks.cpp:10-21")
sumsq = CatalogueFn("sumsq", "sum(x.^2)")
prod = CatalogueFn("prod", "prod(x)")
tanh_affine = CatalogueFn("tanh_affine", "f(x) = tanh.(A*x .+ b); params A (m x n), b (m)")
brusselator = CatalogueFn("brusselator", "f!(dx, x): 1-D Brusselator RHS, x = [u; v]; param alpha", lambda n: n)
mlp2 = CatalogueFn("mlp2", "f(x) = tanh.(A2 * tanh.(A1*x .+ b1) .+ b2); params A1 (h x n), b1 (h), A2 (m x h), b2 (m)")
jactest = CatalogueFn("jactest", "test/JacobianTest.jl:23-31 target (n=3, m=4)", lambda n: 4)


# ---- generic callables on B200DualArray ------------------------------------------
def rosenbrock_generic(x):
    """sum((1 .- x[1:end-1]).^2 .+ 100 .* (x[2:end] .- x[1:end-1].^2).^2)"""
    xi, xn = x[:-1], x[1:]
    return ((1.0 - xi) ** 2 + 100.0 * (xn - xi ** 2) ** 2).sum()


def ackley_generic(x):
    a, b, c = 20.0, -0.2, 2.0 * math.pi
    len_recip = 1.0 / len(x)
    sum_sqrs = (x ** 2).sum()
    sum_cos = A.cos(c * x).sum()
    return (-a) * A.exp(b * A.sqrt(len_recip * sum_sqrs)) - A.exp(len_recip * sum_cos) + a + math.e


def sumsq_generic(x):
    return (x ** 2).sum()


def prod_generic(x):
    return x.prod()
