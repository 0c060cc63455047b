"""Reference evaluator of op-programs (include/fdb200.h, instruction kinds 0-11) in NumPy -- test infrastructure.

Evaluates a traced `fdb200.trace.Program` the slow, obvious way: every register is a full forward-mode dual number
(value, gradient with respect to ALL n inputs) carried as arrays over the term index, with the derivative rules of
SURVEY.md Appendix A (src/dual.jl).  No chunking, no lane sharing, no kernels: it pins what a program MEANS -- guards,
stage markers, parameter arrays, accumulator numbering -- so the tracer and the instruction encoding are checked on the
CPU, independently of the CUDA interpreter and of the generated functors.
"""
import numpy as np

from fdb200._lib import OP1, OP2

K_LOADX, K_UNARY, K_DD, K_DC, K_CD, K_ACC, K_MOV, K_DP, K_PD, K_GUARD, K_STAGE, K_LOADS = range(12)
_U = {v: k for k, v in OP1.items()}
_B = {v: k for k, v in OP2.items()}


class D:
    """value v (shape (m,)) and gradient g (shape (m, n)) for m terms at once"""

    def __init__(self, v, g):
        self.v, self.g = np.asarray(v, dtype=np.float64), np.asarray(g, dtype=np.float64)

    @staticmethod
    def const(c, m, n):
        return D(np.full(m, float(c)), np.zeros((m, n)))

    def chain(self, val, deriv):                                  # Dual(val, deriv * partials), src/dual.jl:213-215
        return D(val, deriv[:, None] * self.g)


def unary(op, x):
    v = x.v
    if op == "neg": return D(-v, -x.g)
    if op == "exp": e = np.exp(v); return x.chain(e, e)
    if op == "log": return x.chain(np.log(v), 1.0 / v)
    if op == "sin": return x.chain(np.sin(v), np.cos(v))
    if op == "cos": return x.chain(np.cos(v), -np.sin(v))
    if op == "tanh": t = np.tanh(v); return x.chain(t, 1.0 - t * t)
    if op == "sqrt": r = np.sqrt(v); return x.chain(r, 1.0 / (2.0 * r))
    if op == "abs": return x.chain(np.abs(v), np.where(np.signbit(v), -1.0, 1.0))
    if op == "abs2": return x.chain(v * v, v + v)
    if op == "inv": i = 1.0 / v; return x.chain(i, -(i * i))
    if op == "pow0": return D(np.ones_like(v), np.zeros_like(x.g))
    if op == "pow1": return D(v, x.g)
    if op == "pow2": return x.chain(v * v, 2.0 * v)
    if op == "pow3": return x.chain(v * v * v, 3.0 * (v * v))
    if op == "tan": t = np.tan(v); return x.chain(t, 1.0 + t * t)
    if op == "exp2": e = np.exp2(v); return x.chain(e, e * np.log(2.0))
    if op == "log2": return x.chain(np.log2(v), (1.0 / v) / np.log(2.0))
    if op == "log10": return x.chain(np.log10(v), (1.0 / v) / np.log(10.0))
    if op == "log1p": return x.chain(np.log1p(v), 1.0 / (v + 1.0))
    if op == "expm1": return x.chain(np.expm1(v), np.exp(v))
    if op == "sinh": return x.chain(np.sinh(v), np.cosh(v))
    if op == "cosh": return x.chain(np.cosh(v), np.sinh(v))
    if op == "asin": return x.chain(np.arcsin(v), 1.0 / np.sqrt(1.0 - v * v))
    if op == "acos": return x.chain(np.arccos(v), -1.0 / np.sqrt(1.0 - v * v))
    if op == "atan": return x.chain(np.arctan(v), 1.0 / (1.0 + v * v))
    if op == "cbrt": c = np.cbrt(v); return x.chain(c, 1.0 / (3.0 * c * c))
    raise ValueError(op)


def binary(op, a, b):
    if op == "add": return D(a.v + b.v, a.g + b.g)
    if op == "sub": return D(a.v - b.v, a.g - b.g)
    if op == "mul": return D(a.v * b.v, a.g * b.v[:, None] + b.g * a.v[:, None])
    if op == "div": return D(a.v / b.v, a.g / b.v[:, None] - b.g * (a.v / (b.v * b.v))[:, None])
    if op == "pow":
        e = a.v ** b.v
        with np.errstate(divide="ignore", invalid="ignore"):
            lg = np.where(np.all(b.g == 0.0, axis=1), 1.0, e * np.log(a.v))
        return D(e, a.g * (b.v * a.v ** (b.v - 1.0))[:, None] + b.g * lg[:, None])
    if op == "max": pick = a.v > b.v; return D(np.where(pick, a.v, b.v), np.where(pick[:, None], a.g, b.g))
    if op == "min": pick = a.v < b.v; return D(np.where(pick, a.v, b.v), np.where(pick[:, None], a.g, b.g))
    if op == "atan2":
        h = a.v * a.v + b.v * b.v
        return D(np.arctan2(a.v, b.v), a.g * (b.v / h)[:, None] - b.g * (a.v / h)[:, None])
    if op == "hypot":
        h = np.hypot(a.v, b.v)
        return D(h, a.g * (a.v / h)[:, None] + b.g * (b.v / h)[:, None])
    raise ValueError(op)


def _run(instrs, R, consts, m, n, xload=None, params=None, acc=None, S=None, idx=None):
    """Execute a list of instructions on the register file R (dict reg -> D) for the term indices `idx`."""
    pc = 0
    while pc < len(instrs):
        op, dst, a, b = (int(t) for t in instrs[pc])
        kind, code = op >> 6, op & 63
        if kind == K_LOADX: R[dst] = xload(b)
        elif kind == K_UNARY: R[dst] = unary(_U[code], R[a])
        elif kind == K_DD: R[dst] = binary(_B[code], R[a], R[b])
        elif kind == K_DC: R[dst] = binary(_B[code], R[a], D.const(consts[b], m, n))
        elif kind == K_CD: R[dst] = binary(_B[code], D.const(consts[a], m, n), R[b])
        elif kind == K_DP: R[dst] = binary(_B[code], R[a], D(params[b][idx], np.zeros((m, n))))
        elif kind == K_PD: R[dst] = binary(_B[code], D(params[a][idx], np.zeros((m, n))), R[b])
        elif kind == K_ACC: (_acc_mul if b == 1 else _acc_add)(acc, dst, R[a])
        elif kind == K_MOV: R[dst] = R[a]
        elif kind == K_LOADS: R[dst] = D(np.repeat(S[a].v, m), np.repeat(S[a].g, m, axis=0))
        else: raise ValueError(f"kind {kind} not executable here")
        pc += 1


def _acc_add(acc, k, d):
    v, g = d.v.sum(keepdims=True), d.g.sum(axis=0, keepdims=True)
    if k in acc:
        acc[k] = D(acc[k].v + v, acc[k].g + g)
    else:
        acc[k] = D(v, g)
    return acc[k]


def _acc_mul(acc, k, d):
    """product accumulator (`prod`, PK_ACC with b = 1): starts at one(Dual), combines by Dual*Dual term by term"""
    n = d.g.shape[1]
    cur = acc.get(k, D(np.ones(1), np.zeros((1, n))))
    v, g = cur.v.copy(), cur.g.copy()
    for i in range(d.v.size):
        g = g * d.v[i] + d.g[i:i + 1] * v
        v = v * d.v[i]
    acc[k] = D(v, g)
    return acc[k]


def _term_groups(term, n, halo):
    """Split a term instruction list into (instructions, term count) groups following the GUARD pseudo-instructions."""
    groups, pc, plain = [], 0, []
    while pc < len(term):
        op, _, a, b = (int(t) for t in term[pc])
        if (op >> 6) == K_GUARD:
            groups.append((term[pc + 1: pc + 1 + b], n - a))
            pc += 1 + b
        else:
            plain.append(term[pc]); pc += 1
    if plain:
        groups.append((plain, n - halo))
    return groups


def evaluate(prog, x):
    """-> (f(x), gradient) of a scalar-valued Program; for an array-valued one -> (y, Jacobian)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    term = [tuple(int(t) for t in ins) for ins in prog.term]
    epi = [tuple(int(t) for t in ins) for ins in prog.epi]
    consts, params = prog.consts, prog.params
    eye = np.eye(n)

    def loader(idx):
        return lambda off: D(x[idx + off], eye[idx + off])

    stages = [i for i, ins in enumerate(term) if (ins[0] >> 6) == K_STAGE]
    acc = {}
    if not stages:
        if prog.n_acc == 0:                                      # array form: y[i] = result register after the term part
            m = n - prog.halo; idx = np.arange(m); R = {}
            _run(term, R, consts, m, n, loader(idx), params, acc, None, idx)
            return R[prog.result_reg].v, R[prog.result_reg].g
        for instrs, m in _term_groups(term, n, prog.halo):
            idx = np.arange(max(m, 0)); R = {}
            _run(instrs, R, consts, idx.size, n, loader(idx), params, acc, None, idx)
        prod = {int(ins[1]) for ins in term if (ins[0] >> 6) == K_ACC and ins[3] == 1}
        R = {k: acc.get(k, D.const(1.0 if k in prod else 0.0, 1, n)) for k in range(prog.n_acc)}
        _run(epi, R, consts, 1, n)
        y = R[prog.result_reg]
        return float(y.v[0]), y.g[0]
    st1, st2 = stages
    m = n - prog.halo; idx = np.arange(m); R = {}
    _run(term[:st1], R, consts, m, n, loader(idx), params, acc, None, idx)
    k1 = 1 + max((ins[1] for ins in term[:st1] if (ins[0] >> 6) == K_ACC), default=-1)
    S = {k: acc.get(k, D.const(0.0, 1, n)) for k in range(k1)}
    _run(term[st1 + 1: st2], S, consts, 1, n)                    # first-stage epilogue; its register file is S
    R = {}
    _run(term[st2 + 1:], R, consts, m, n, loader(idx), params, acc, S, idx)
    if prog.out_len > 1 and not epi:                             # array-valued two-pass form: y[i] = result register of the second broadcast
        return R[prog.result_reg].v, R[prog.result_reg].g
    R = {k: acc.get(k, D.const(0.0, 1, n)) for k in range(k1, prog.n_acc)}
    _run(epi, R, consts, 1, n, S=S)
    y = R[prog.result_reg]
    return float(y.v[0]), y.g[0]
