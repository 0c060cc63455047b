"""Flatten a user function into an op-program (include/fdb200.h, fdb_program_*).

The Julia shim would do this by walking the `Broadcasted` tree of `f`'s broadcast expressions
(a custom BroadcastStyle); here the same flattening is done by calling `f` once on a symbolic array
whose operators record instead of compute.  Supported shape: elementwise src/dual.jl arithmetic on
unit-stride views `x[a:b]` of the input (all of one common length), NumPy arrays the function closes over
(data: `w * x**2` -- they become array parameters read at the term index, with the Real (x) Dual rules), `sum()`
reductions of such arrays, scalar Dual arithmetic on the reduced values, and ONE re-entry of reduced values into a
second broadcast (`sum((x .- mean(x)).^2)`: a two-pass program -- stage levels term -> scalar -> term2 -> scalar2).  Anything else raises TraceError and the caller
falls back to the reference chunk loop on materialised device duals.
"""
import numpy as np

from ._lib import OP1, OP2

K_LOADX, K_UNARY, K_DD, K_DC, K_CD, K_ACC, K_MOV, K_DP, K_PD, K_GUARD, K_STAGE, K_LOADS = range(12)
# limits of the program representation (what the run-time specialised kernels accept) and of the per-thread interpreter
MAX_INSTR, MAX_CONST, MAX_REG, MAX_ACC, MAX_PARAMS = 256, 64, 32, 8, 4
I_MAX_INSTR, I_MAX_CONST, I_MAX_REG, I_MAX_ACC = 96, 24, 12, 4


class TraceError(Exception):
    pass


_NUM = (int, float, np.floating, np.integer)


class _Node:
    __slots__ = ("kind", "op", "args", "const", "stage", "id", "is_z")

    def __init__(self, tr, kind, op=0, args=(), const=None, stage="term"):
        self.kind, self.op, self.args, self.const, self.stage = kind, op, args, const, stage
        self.is_z = False
        self.id = len(tr.nodes)
        tr.nodes.append(self)


_LEVEL = {"term": 0, "scalar": 1, "term2": 2, "scalar2": 3}


class _Sym:
    """Common operator plumbing of the symbolic array (stages 'term', 'term2') and symbolic scalar ('scalar', 'scalar2')."""
    __array_ufunc__ = None          # `ndarray * sym` defers to sym.__rmul__ instead of broadcasting over an object

    def __init__(self, tr, node, length):
        self.tr, self.node, self.n = tr, node, length

    def _like(self, node):
        return type(self)(self.tr, node, self.n)

    def _check(self, other):
        if not isinstance(other, _Sym) or other.tr is not self.tr:
            raise TraceError("operand is not part of this trace")
        if type(other) is not type(self):
            raise TraceError("operand kinds differ")
        if other.n != self.n:
            from ._lib import DimensionMismatch
            raise DimensionMismatch(f"arrays could not be broadcast to a common size ({self.n} vs {other.n})")

    def _map1(self, op):
        return self._like(_Node(self.tr, K_UNARY, OP1[op], (self.node,), stage=self.node.stage))

    _trace_unary = _map1

    def _map2(self, op, other, reverse=False):
        code = OP2[op]
        if isinstance(other, _NUM):
            if code > OP2["pow"]:
                raise TraceError(f"{op} needs two Dual operands")
            kind = K_CD if reverse else K_DC
            return self._like(_Node(self.tr, kind, code, (self.node,), const=float(other), stage=self.node.stage))
        if isinstance(other, np.ndarray):
            return self._map2_param(code, op, other, reverse)
        if isinstance(self, TScalar) and isinstance(other, TArray):
            return other._map2(op, self, not reverse)
        if isinstance(self, TArray) and isinstance(other, TScalar) and other.tr is self.tr:
            # a reduced value re-enters a broadcast: Vector{Dual} .(op) Dual -> second pass
            if _LEVEL[other.node.stage] != 1:
                raise TraceError("only one re-entry of reduced values into a broadcast (two passes) is supported")
            a, b = (other.node, self.node) if reverse else (self.node, other.node)
            return self._like(_Node(self.tr, K_DD, code, (a, b), stage="term2"))
        self._check(other)
        a, b = (other.node, self.node) if reverse else (self.node, other.node)
        lvl = max(_LEVEL[a.stage], _LEVEL[b.stage])
        stage = {0: "term", 1: "scalar", 2: "term2", 3: "scalar2"}[lvl]
        return self._like(_Node(self.tr, K_DD, code, (a, b), stage=stage))

    def _map2_param(self, code, op, arr, reverse):
        """A closed-over data array: Dual (x) Real elementwise, the array read at the term index."""
        if not isinstance(self, TArray) or code > OP2["pow"]:
            raise TraceError(f"{op} with a data array is not expressible")
        if arr.ndim != 1 or arr.size != self.n:
            from ._lib import DimensionMismatch
            raise DimensionMismatch(f"arrays could not be broadcast to a common size ({self.n} vs {arr.shape})")
        tr = self.tr
        for k, a in enumerate(tr.params):
            if a is arr:
                break
        else:
            if len(tr.params) >= MAX_PARAMS:
                raise TraceError("too many data arrays")
            tr.params.append(arr)
            k = len(tr.params) - 1
        return self._like(_Node(tr, K_PD if reverse else K_DP, code, (self.node,), const=k, stage=self.node.stage))

    def __add__(self, o): return self._map2("add", o)
    def __radd__(self, o): return self._map2("add", o, True)
    def __sub__(self, o): return self._map2("sub", o)
    def __rsub__(self, o): return self._map2("sub", o, True)
    def __mul__(self, o): return self._map2("mul", o)
    def __rmul__(self, o): return self._map2("mul", o, True)
    def __truediv__(self, o): return self._map2("div", o)
    def __rtruediv__(self, o): return self._map2("div", o, True)
    def __neg__(self): return self._map1("neg")

    def __pow__(self, o):
        if isinstance(o, int) and not isinstance(o, bool) and 0 <= o <= 3:      # Base.literal_pow, src/dual.jl:587-597
            return self._map1(f"pow{o}")
        return self._map2("pow", o)

    def __rpow__(self, o): return self._map2("pow", o, True)


class TScalar(_Sym):
    """A reduced Dual (result of sum) or scalar Dual math on such values: the epilogue stage."""


class TArray(_Sym):
    def __len__(self):
        return self.n

    def __getitem__(self, sl):
        if self.node.kind != K_LOADX or self.node.args:
            raise TraceError("views are only supported directly on the input array")
        if not isinstance(sl, slice):
            raise TraceError("scalar indexing")
        start, stop, step = sl.indices(self.n)
        if step != 1:
            raise TraceError("non-unit stride")
        node = _Node(self.tr, K_LOADX, 0, (), const=self.node.const + start)
        return TArray(self.tr, node, max(stop - start, 0))

    def sum(self):
        tr = self.tr
        if len(tr.accs) >= MAX_ACC:
            raise TraceError("too many reductions")
        tr.accs.append(self.node)
        tr.acc_prod.append(False)
        tr.acc_lengths.append(self.n)
        tr.term_lengths.add(self.n)
        node = _Node(tr, K_ACC, len(tr.accs) - 1, (), stage="scalar" if _LEVEL[self.node.stage] == 0 else "scalar2")
        tr.acc_nodes.append(node)
        return TScalar(tr, node, 1)

    def mean(self):
        """Statistics.mean: sum(x) / length(x) (Dual / Int, src/dual.jl:538)."""
        return self.sum() / self.n

    def dot(self, other):
        """LinearAlgebra.dot of two real vectors: sum of elementwise products."""
        return (self * other).sum()

    def prod(self):
        """prod(x): an accumulator that starts at one(Dual) and combines by Dual*Dual (specialised kernels only)."""
        if _LEVEL[self.node.stage] != 0:
            raise TraceError("prod over a second-pass broadcast")
        s = self.sum()
        self.tr.acc_prod[-1] = True
        return s

    def __rmatmul__(self, A):
        """A @ x for a closed-over Float64 matrix: `A * xdual` of the reference (LinearAlgebra.generic_matvecmul! over Dual,
        src/dual.jl:476-490 for the Real*Dual / Dual+Dual it is made of).  Only the input array itself may be contracted,
        once; what follows is an elementwise program over z = A*x, whose Dual z_i the device builds from the FP64
        tensor-core contraction (fdb_program_matvec_jacobian_chunked)."""
        tr = self.tr
        if self.node.kind != K_LOADX or self.node.args or self.node.const != 0 or self.n != tr.n or self.node.stage != "term" \
                or getattr(self.node, "is_z", False):
            raise TraceError("A @ v is only supported for v = the whole input array")
        if tr.matvec is not None:
            raise TraceError("only one matrix-vector product per function")
        A = np.asarray(A)
        if A.ndim != 2 or A.dtype != np.float64:
            raise TraceError("A @ x needs a 2-d float64 matrix")
        if A.shape[1] != self.n:
            from ._lib import DimensionMismatch
            raise DimensionMismatch(f"A has dimensions {A.shape}, vector has length {self.n}")
        tr.matvec = A
        tr.x_node = self.node
        node = _Node(tr, K_LOADX, 0, (), const=0)
        node.is_z = True
        tr.z_node = node
        return TArray(tr, node, A.shape[0])


class _Trace:
    def __init__(self, n):
        self.n = n
        self.nodes, self.accs, self.acc_lengths, self.term_lengths, self.params, self.acc_nodes = [], [], [], set(), [], []
        self.matvec, self.x_node, self.z_node = None, None, None
        self.acc_prod = []


def _emit(tr, roots, stage, preloaded, sreg=None):
    """Linearise the DAG reachable from `roots` (restricted to `stage`, a name or a set of names) into instructions with
    register reuse.  `sreg` maps first-stage scalar nodes to their slot in S: such operands are fetched with LOADS."""
    order, seen = [], set()
    stages = {stage} if isinstance(stage, str) else set(stage)
    stage = "/".join(sorted(stages))

    def visit(nd):
        if nd.id in seen:
            return
        if nd.stage not in stages:
            if sreg is not None and nd.id in sreg:          # value published by the first-stage epilogue
                seen.add(nd.id)
                order.append(nd)
            return
        seen.add(nd.id)
        for a in nd.args:
            visit(a)
        order.append(nd)

    for r in roots:
        visit(r)
    last_use = {}
    for pos, nd in enumerate(order):
        for a in nd.args:
            last_use[a.id] = pos
    for r in roots:
        last_use[r.id] = len(order) + 1
    consts, instrs, reg = tr.consts, [], dict(preloaded)
    free = [r for r in range(MAX_REG - 1, -1, -1) if r not in reg.values()]

    def cidx(v):
        for i, c in enumerate(consts):
            if c == v and np.signbit(c) == np.signbit(v):
                return i
        if len(consts) >= MAX_CONST:
            raise TraceError("too many constants")
        consts.append(v)
        return len(consts) - 1

    for pos, nd in enumerate(order):
        if nd.id in reg:                         # preloaded accumulator register
            continue
        foreign = nd.stage not in stages
        srcs = [] if foreign else [reg[a.id] for a in nd.args]
        for a in (() if foreign else nd.args):   # release operands whose last use is this instruction
            if last_use.get(a.id) == pos and a.id in reg and reg[a.id] not in free and a.id not in preloaded:
                free.append(reg[a.id])
        if not free:
            raise TraceError("expression needs more than %d live Dual registers" % MAX_REG)
        dst = free.pop()
        reg[nd.id] = dst
        if nd.stage not in stages:
            instrs.append((K_LOADS << 6, dst, sreg[nd.id], 0))
        elif nd.kind == K_LOADX:
            instrs.append((K_LOADX << 6, dst, 0, int(nd.const)))
        elif nd.kind == K_UNARY:
            instrs.append(((K_UNARY << 6) | nd.op, dst, srcs[0], 0))
        elif nd.kind == K_DD:
            instrs.append(((K_DD << 6) | nd.op, dst, srcs[0], srcs[1]))
        elif nd.kind == K_DC:
            instrs.append(((K_DC << 6) | nd.op, dst, srcs[0], cidx(nd.const)))
        elif nd.kind == K_CD:
            instrs.append(((K_CD << 6) | nd.op, dst, cidx(nd.const), srcs[0]))
        elif nd.kind == K_DP:
            instrs.append(((K_DP << 6) | nd.op, dst, srcs[0], int(nd.const)))
        elif nd.kind == K_PD:
            instrs.append(((K_PD << 6) | nd.op, dst, int(nd.const), srcs[0]))
        else:
            raise TraceError("unexpected node in stage " + stage)
        if len(instrs) > MAX_INSTR:
            raise TraceError("program too long")
    return instrs, reg


class Program:
    """term/epilogue instruction arrays + constants, ready for fdb_program_*_chunked."""

    def __init__(self, term, epi, consts, n_acc, halo, result_reg, out_len, params=(), matvec=None):
        self.term = np.asarray(term, dtype=np.int32).reshape(-1, 4)
        self.epi = np.asarray(epi, dtype=np.int32).reshape(-1, 4)
        self.consts = np.asarray(consts, dtype=np.float64)
        self.n_acc, self.halo, self.result_reg, self.out_len = n_acc, halo, result_reg, out_len
        self.matvec = matvec                     # A of `A @ x`: the program is elementwise over z = A*x (LOADX 0 = z_i)
        self.params = [np.ascontiguousarray(a, dtype=np.float64) for a in params]     # data arrays, one entry per term
        # data arrays and guarded groups exist only in the run-time specialised kernels, not in the interpreter
        kinds = (self.term[:, 0] >> 6) if self.term.size else np.zeros(0, dtype=np.int32)
        self.two_pass = bool(np.any(kinds == K_STAGE))
        self.has_prod = bool(np.any((kinds == K_ACC) & (self.term[:, 3] == 1))) if self.term.size else False
        regs = [int(r) for part in (self.term, self.epi) for ins in part for r in ((ins[1],) if (ins[0] >> 6) not in (K_ACC, K_GUARD, K_STAGE) else ())]
        self.beyond_interpreter = (len(self.term) > I_MAX_INSTR or len(self.epi) > I_MAX_INSTR or self.consts.size > I_MAX_CONST or
                                   n_acc > I_MAX_ACC or max(regs + [result_reg]) >= I_MAX_REG)
        self.needs_jit = bool(self.params) or bool(np.any(kinds == K_GUARD)) or self.two_pass or self.has_prod or self.beyond_interpreter


def _common_halo(tr, lengths):
    if len(lengths) != 1:
        raise TraceError("reductions over arrays of different lengths")
    L = lengths.pop()
    halo = tr.n - L
    offs = [nd.const for nd in tr.nodes if nd.kind == K_LOADX and nd.stage == "term"]
    if halo < 0 or any(o > halo for o in offs):
        raise TraceError("view outside the input")
    return L, halo


def trace_scalar(f, n):
    """f(x) -> scalar: returns a Program for fdb_program_gradient_chunked, or raises TraceError."""
    tr = _Trace(n)
    tr.consts = []
    x = TArray(tr, _Node(tr, K_LOADX, 0, (), const=0), n)
    y = f(x)
    if not isinstance(y, TScalar):
        raise TraceError("f did not return a reduced scalar")
    if tr.matvec is not None:
        raise TraceError("a reduction over A @ x has no batched kernel (the reference loop contracts on the tensor cores)")
    if any(_LEVEL[nd.stage] >= 2 for nd in tr.accs):
        if any(tr.acc_prod):
            raise TraceError("prod in a two-pass function")
        return _two_pass_program(tr, y)
    if len(tr.term_lengths) == 1:
        L, halo = _common_halo(tr, set(tr.term_lengths))
        # term part: evaluate every accumulated expression, then ACC
        term, reg = _emit(tr, tr.accs, "term", {})
        for k, nd in enumerate(tr.accs):
            term.append((K_ACC << 6, k, reg[nd.id], int(tr.acc_prod[k])))
    else:
        # reductions over views of different lengths (sum(g.(x[2:end], x[1:end-1])) + sum(h.(x))): one guarded group per
        # reduction, term i of group k runs while i + halo_k < n (specialised kernels only)
        if not tr.accs or min(tr.acc_lengths) < 0 or max(tr.acc_lengths) > tr.n:
            raise TraceError("view outside the input")
        term, halo = [], tr.n - min(tr.acc_lengths)
        for k, (nd, L) in enumerate(zip(tr.accs, tr.acc_lengths)):
            group, reg = _emit(tr, [nd], "term", {})
            if any((ins[0] >> 6) == K_LOADX and ins[3] > tr.n - L for ins in group):
                raise TraceError("view outside the input")
            group.append((K_ACC << 6, k, reg[nd.id], int(tr.acc_prod[k])))
            term.append((K_GUARD << 6, 0, tr.n - L, len(group)))
            term += group
    if len(term) > MAX_INSTR:
        raise TraceError("program too long")
    pre = {nd.id: nd.op for nd in tr.nodes if nd.kind == K_ACC}
    epi, ereg = _emit(tr, [y.node], "scalar", pre)
    return Program(term, epi, tr.consts, len(tr.accs), halo, ereg[y.node.id], 1, tr.params)


def _two_pass_program(tr, y):
    """term1 | STAGE 1 | epilogue1 | STAGE 2 | term2 in the term part, the final epilogue in the epilogue part.
    Pass-1 sums take the accumulator indices [0, K1), pass-2 sums [K1, K)."""
    L, halo = _common_halo(tr, set(tr.term_lengths))
    first = [k for k, nd in enumerate(tr.accs) if _LEVEL[nd.stage] == 0]
    second = [k for k, nd in enumerate(tr.accs) if _LEVEL[nd.stage] == 2]
    for new, old in enumerate(first + second):                     # renumber: pass-1 accumulators first
        tr.acc_nodes[old].op = new
    K1 = len(first)
    term, reg = _emit(tr, [tr.accs[k] for k in first], "term", {})
    for k in first:
        term.append((K_ACC << 6, tr.acc_nodes[k].op, reg[tr.accs[k].id], 0))
    # first-stage scalars needed later: operands of second-pass nodes and of the final epilogue
    late = [nd for nd in tr.nodes if _LEVEL[nd.stage] >= 2]
    needed, seen = [], set()
    for nd in late:
        for a in nd.args:
            if _LEVEL[a.stage] == 1 and a.id not in seen:
                seen.add(a.id)
                needed.append(a)
    if _LEVEL[y.node.stage] == 1:
        raise TraceError("the second pass does not contribute to the result")
    pre1 = {tr.acc_nodes[k].id: tr.acc_nodes[k].op for k in first}
    epi1, sreg = _emit(tr, needed, "scalar", pre1)
    sreg = {nd.id: sreg[nd.id] for nd in needed}
    term.append((K_STAGE << 6, 0, 1, 0))
    term += epi1
    term.append((K_STAGE << 6, 0, 2, 0))
    term2, reg2 = _emit(tr, [tr.accs[k] for k in second], {"term", "term2"}, {}, sreg)
    term += term2
    for k in second:
        term.append((K_ACC << 6, tr.acc_nodes[k].op, reg2[tr.accs[k].id], 0))
    if len(term) > MAX_INSTR:
        raise TraceError("program too long")
    pre2 = {tr.acc_nodes[k].id: tr.acc_nodes[k].op for k in second}
    epi, ereg = _emit(tr, [y.node], "scalar2", pre2, sreg)
    return Program(term, epi, tr.consts, len(tr.accs), halo, ereg[y.node.id], 1, tr.params)


def trace_array(f, n):
    """f(x) -> array (elementwise / stencil): Program for fdb_program_jacobian_chunked."""
    tr = _Trace(n)
    tr.consts = []
    x = TArray(tr, _Node(tr, K_LOADX, 0, (), const=0), n)
    y = f(x)
    if not isinstance(y, TArray):
        raise TraceError("f did not return an array")
    if tr.accs:
        if tr.matvec is not None:
            raise TraceError("reductions over A @ x inside an array-valued f")
        return _two_pass_array_program(tr, y)
    if tr.matvec is not None:
        # elementwise over z = A @ x: every leaf must be z itself (x may not appear again, nor views of z)
        m = tr.matvec.shape[0]
        if y.n != m:
            raise TraceError("result length differs from the rows of A")
        term, reg = _emit(tr, [y.node], "term", {})
        used = {ins[3] for ins in term if (ins[0] >> 6) == K_LOADX}
        reach = _reachable(y.node)
        if tr.x_node.id in reach or any(nd.kind == K_LOADX and not nd.is_z for nd in tr.nodes if nd.id in reach) or used - {0}:
            raise TraceError("x may only enter through A @ x")
        return Program(term, [], tr.consts, 0, 0, reg[y.node.id], m, tr.params, matvec=tr.matvec)
    L, halo = _common_halo(tr, {y.n})
    term, reg = _emit(tr, [y.node], "term", {})
    return Program(term, [], tr.consts, 0, halo, reg[y.node.id], L, tr.params)


def _two_pass_array_program(tr, y):
    """y = g.(x, reduced values): term1 | STAGE 1 | epilogue1 | STAGE 2 | term2, the result register of term2 is y[i]
    (softmax: exp.(x) ./ sum(exp.(x))).  Specialised kernels only (jacobian_sweep2_body)."""
    if any(tr.acc_prod) or any(_LEVEL[nd.stage] != 0 for nd in tr.accs):
        raise TraceError("only plain first-pass sums may feed an array-valued second broadcast")
    if _LEVEL[y.node.stage] != 2:
        raise TraceError("the reductions do not contribute to the result")
    L, halo = _common_halo(tr, set(tr.term_lengths) | {y.n})
    term, reg = _emit(tr, tr.accs, "term", {})
    for k, nd in enumerate(tr.accs):
        term.append((K_ACC << 6, k, reg[nd.id], 0))
    needed, seen = [], set()
    for nd in tr.nodes:
        if _LEVEL[nd.stage] >= 2:
            for a in nd.args:
                if _LEVEL[a.stage] == 1 and a.id not in seen:
                    seen.add(a.id)
                    needed.append(a)
    pre1 = {nd.id: nd.op for nd in tr.acc_nodes}
    epi1, sreg = _emit(tr, needed, "scalar", pre1)
    sreg = {nd.id: sreg[nd.id] for nd in needed}
    term.append((K_STAGE << 6, 0, 1, 0))
    term += epi1
    term.append((K_STAGE << 6, 0, 2, 0))
    term2, reg2 = _emit(tr, [y.node], {"term", "term2"}, {}, sreg)
    term += term2
    if len(term) > MAX_INSTR:
        raise TraceError("program too long")
    return Program(term, [], tr.consts, len(tr.accs), halo, reg2[y.node.id], L, tr.params)


def _reachable(root):
    seen, stack = set(), [root]
    while stack:
        nd = stack.pop()
        if nd.id in seen:
            continue
        seen.add(nd.id)
        stack.extend(nd.args)
    return seen
