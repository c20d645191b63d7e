"""Tracing of stimulus closures f(x, t) into the postfix programs of tb_assemble_source_program.

The reference evaluates an `AnalyticalCoefficient`'s closure at every quadrature point inside the element loop
(src/modeling/core/analytical_coefficient.jl:80-101).  A closure cannot cross the C ABI, its expression can: `trace_source`
calls f ONCE with tracing numbers, linearises the recorded expression into postfix code (include/tbolt_b200.h,
TB_SRC_PROGRAM) and checks the program against f itself at a few sample points.  What cannot be traced -- Python control
flow on x or t (`if`, `and`, `max(...)` go through `bool()`), math.* functions, more than 96 instructions -- returns None and
the caller keeps the host-evaluated path (tb_assemble_source_qp).  Branch-free forms are available as `where`, `minimum`,
`maximum`, `norm`, `&`, `|`, `~`; closures written with plain numpy trace as well: ufuncs on a tracing number
(np.maximum(x[0], 1.0), np.exp(-r), np.less(t, 2.0), np.power(x[1], 2) ...) dispatch through `__array_ufunc__`, reductions
over the coordinate vector (np.linalg.norm(x), x @ x) through numpy's object loops.

The Julia binding does the same with a `TracedReal <: Real` number type.
"""
from __future__ import annotations

import math

import numpy as np

OPS = ["X", "T", "CONST", "ADD", "SUB", "MUL", "DIV", "MIN", "MAX", "POW", "LT", "LE", "GT", "GE", "EQ", "NE", "AND", "OR",
       "NEG", "ABS", "SQRT", "EXP", "LOG", "SIN", "COS", "TANH", "NOT", "SELECT"]
OP = {name: i for i, name in enumerate(OPS)}
MAXCODE, MAXCONST, MAXSTACK = 96, 24, 16


class TraceError(Exception):
    pass


class Sym:
    """A tracing number: every operation returns a new node of the expression tree."""
    __slots__ = ("op", "args")
    __array_priority__ = 1000     # numpy defers to our reflected operators
    __hash__ = object.__hash__

    def __init__(self, op, *args):
        self.op, self.args = op, args

    @staticmethod
    def lift(v):
        if isinstance(v, Sym):
            return v
        if isinstance(v, (bool, np.bool_)):
            return Sym("CONST", 1.0 if v else 0.0)
        if isinstance(v, (int, float, np.integer, np.floating)):
            return Sym("CONST", float(v))
        if isinstance(v, np.ndarray) and v.ndim == 0 and v.dtype != object:
            return Sym.lift(v.item())
        raise TraceError(f"cannot trace a value of type {type(v).__name__}")

    def _bin(self, op, other, swap=False):
        o = Sym.lift(other)
        return Sym(op, o, self) if swap else Sym(op, self, o)

    def __add__(self, o): return self._bin("ADD", o)
    def __radd__(self, o): return self._bin("ADD", o, True)
    def __sub__(self, o): return self._bin("SUB", o)
    def __rsub__(self, o): return self._bin("SUB", o, True)
    def __mul__(self, o): return self._bin("MUL", o)
    def __rmul__(self, o): return self._bin("MUL", o, True)
    def __truediv__(self, o): return self._bin("DIV", o)
    def __rtruediv__(self, o): return self._bin("DIV", o, True)
    def __lt__(self, o): return self._bin("LT", o)
    def __le__(self, o): return self._bin("LE", o)
    def __gt__(self, o): return self._bin("GT", o)
    def __ge__(self, o): return self._bin("GE", o)
    def __eq__(self, o): return self._bin("EQ", o)
    def __ne__(self, o): return self._bin("NE", o)
    def __and__(self, o): return self._bin("AND", o)
    def __rand__(self, o): return self._bin("AND", o, True)
    def __or__(self, o): return self._bin("OR", o)
    def __ror__(self, o): return self._bin("OR", o, True)
    def __invert__(self): return Sym("NOT", self)
    def __neg__(self): return Sym("NEG", self)
    def __pos__(self): return self
    def __abs__(self): return Sym("ABS", self)

    def __pow__(self, o):
        # Julia lowers x^2 and x^3 with literal exponents to x*x and x*x*x (Base.literal_pow); keep those exact
        if isinstance(o, (int, np.integer)) and not isinstance(o, bool) and 1 <= int(o) <= 3:
            r = self
            for _ in range(int(o) - 1):
                r = r * self
            return r
        return self._bin("POW", o)

    def __rpow__(self, o): return self._bin("POW", o, True)

    def __bool__(self):
        raise TraceError("the closure branches on x or t (if / and / or / max()); use where(), minimum(), maximum(), &, |")

    # numpy ufuncs applied to a tracing number (np.maximum(x[0], 1.0), np.exp(-r), np.less(t, 2.0) ...) land here
    _UFUNCS = {"add": "ADD", "subtract": "SUB", "multiply": "MUL", "true_divide": "DIV", "divide": "DIV", "maximum": "MAX",
               "minimum": "MIN", "fmax": "MAX", "fmin": "MIN", "less": "LT", "less_equal": "LE", "greater": "GT",
               "greater_equal": "GE", "equal": "EQ", "not_equal": "NE", "logical_and": "AND", "logical_or": "OR",
               "negative": "NEG", "absolute": "ABS", "fabs": "ABS", "sqrt": "SQRT", "exp": "EXP", "log": "LOG", "sin": "SIN",
               "cos": "COS", "tanh": "TANH", "logical_not": "NOT"}

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        if ufunc.__name__ in ("power", "float_power") and len(inputs) == 2 and inputs[0] is self:
            return self.__pow__(inputs[1])
        if ufunc.__name__ == "square" and len(inputs) == 1:
            return self * self
        op = Sym._UFUNCS.get(ufunc.__name__)
        if op is None:
            raise TraceError(f"numpy.{ufunc.__name__} cannot be traced")
        return Sym(op, *(Sym.lift(v) for v in inputs))

    # numpy ufuncs on object ARRAYS of tracing numbers call the method of the same name, element by element
    def sqrt(self): return Sym("SQRT", self)
    def exp(self): return Sym("EXP", self)
    def log(self): return Sym("LOG", self)
    def sin(self): return Sym("SIN", self)
    def cos(self): return Sym("COS", self)
    def tanh(self): return Sym("TANH", self)
    def conjugate(self): return self
    real = property(lambda self: self)


def _any_sym(*a):
    return any(isinstance(v, Sym) for v in a)


def where(cond, a, b):
    """cond ? a : b without a Python branch (Julia: ifelse)"""
    if _any_sym(cond, a, b):
        return Sym("SELECT", Sym.lift(cond), Sym.lift(a), Sym.lift(b))
    return a if cond else b


def minimum(a, b):
    return Sym("MIN", Sym.lift(a), Sym.lift(b)) if _any_sym(a, b) else (b if b < a else a)


def maximum(a, b):
    return Sym("MAX", Sym.lift(a), Sym.lift(b)) if _any_sym(a, b) else (b if a < b else a)


def norm(x):
    """sqrt(x1^2 + x2^2 + ...), summed left to right"""
    s = x[0] * x[0]
    for v in x[1:]:
        s = s + v * v
    return s.sqrt() if isinstance(s, Sym) else math.sqrt(s)


def _unary(name, fn):
    def g(a):
        return Sym(name, a) if isinstance(a, Sym) else fn(a)
    g.__name__ = name.lower()
    return g


sqrt, exp, log, sin, cos, tanh = (_unary(n, f) for n, f in (("SQRT", math.sqrt), ("EXP", math.exp), ("LOG", math.log),
                                                            ("SIN", math.sin), ("COS", math.cos), ("TANH", math.tanh)))


class SourceProgram:
    """Postfix code (op | arg << 8 per instruction) + constant table; callable like the closure it came from."""

    def __init__(self, code, consts, dim):
        self.code = np.asarray(code, dtype=np.int32)
        self.consts = np.asarray(consts, dtype=np.float64)
        self.dim = dim

    def __len__(self):
        return int(self.code.size)

    def __call__(self, x, t):
        st = []
        for ins in self.code:
            op, arg = OPS[int(ins) & 0xff], int(ins) >> 8
            if op == "X": st.append(float(x[arg]))
            elif op == "T": st.append(float(t))
            elif op == "CONST": st.append(float(self.consts[arg]))
            elif op == "SELECT":
                b = st.pop(); a = st.pop(); c = st.pop()
                st.append(a if c != 0.0 else b)
            elif OP[op] >= OP["NEG"]:
                a = st.pop()
                with np.errstate(all="ignore"):
                    st.append({"NEG": lambda: -a, "ABS": lambda: abs(a), "SQRT": lambda: float(np.sqrt(a)),
                               "EXP": lambda: float(np.exp(a)), "LOG": lambda: float(np.log(a)), "SIN": lambda: math.sin(a),
                               "COS": lambda: math.cos(a), "TANH": lambda: math.tanh(a),
                               "NOT": lambda: 1.0 if a == 0.0 else 0.0}[op]())
            else:
                b = st.pop(); a = st.pop()
                with np.errstate(all="ignore"):
                    st.append({"ADD": lambda: a + b, "SUB": lambda: a - b, "MUL": lambda: a * b,
                               "DIV": lambda: float(np.float64(a) / np.float64(b)), "MIN": lambda: b if b < a else a,
                               "MAX": lambda: b if a < b else a, "POW": lambda: float(np.power(np.float64(a), np.float64(b))),
                               "LT": lambda: float(a < b), "LE": lambda: float(a <= b), "GT": lambda: float(a > b),
                               "GE": lambda: float(a >= b), "EQ": lambda: float(a == b), "NE": lambda: float(a != b),
                               "AND": lambda: float(a != 0.0 and b != 0.0), "OR": lambda: float(a != 0.0 or b != 0.0)}[op]())
        return st[0]


def compile_expr(root: Sym, dim: int) -> SourceProgram:
    code, consts, cidx = [], [], {}
    depth = maxdepth = 0

    def emit(node):
        nonlocal depth, maxdepth
        if node.op == "X":
            code.append(OP["X"] | (node.args[0] << 8)); depth += 1
        elif node.op == "T":
            code.append(OP["T"]); depth += 1
        elif node.op == "CONST":
            v = node.args[0]
            key = np.float64(v).tobytes()
            if key not in cidx:
                if len(consts) >= MAXCONST:
                    raise TraceError("more than %d distinct constants" % MAXCONST)
                cidx[key] = len(consts)
                consts.append(v)
            code.append(OP["CONST"] | (cidx[key] << 8)); depth += 1
        else:
            for a in node.args:
                emit(a)
            code.append(OP[node.op]); depth += 1 - len(node.args)
        maxdepth = max(maxdepth, depth)
        if len(code) > MAXCODE:
            raise TraceError("more than %d instructions" % MAXCODE)

    emit(root)
    if maxdepth > MAXSTACK:
        raise TraceError("expression needs a deeper stack than %d" % MAXSTACK)
    return SourceProgram(code, consts, dim)


def trace_source(f, dim: int, samples=None, rtol=1e-12):
    """f(x, t) -> SourceProgram, or None when f cannot be traced (or the traced program does not reproduce f)."""
    x = np.empty(dim, dtype=object)
    for d in range(dim):
        x[d] = Sym("X", d)
    try:
        import sys
        lim = sys.getrecursionlimit()
        root = Sym.lift(f(x, Sym("T")))
        sys.setrecursionlimit(max(lim, 4000))
        try:
            prog = compile_expr(root, dim)
        finally:
            sys.setrecursionlimit(lim)
    except TraceError:
        return None
    except (TypeError, ValueError, AttributeError, IndexError):
        return None
    # the program must be the closure: compare at sample points (catches closures with captured mutable state, randomness,
    # or branches that were decided at trace time on something other than x and t)
    if samples is None:
        rng = np.random.default_rng(20190)
        samples = [(rng.uniform(-2.0, 2.0, dim), float(tt)) for tt in (0.0, 0.3, 1.7, 2.5)] + [(np.zeros(dim), 0.0)]
    for xs, ts in samples:
        try:
            with np.errstate(all="ignore"):
                want = float(f(np.asarray(xs, dtype=np.float64), ts))
        except Exception:
            return None
        got = prog(xs, ts)
        if not (got == want or abs(got - want) <= rtol * max(abs(want), abs(got)) or (math.isnan(got) and math.isnan(want))):
            return None
    return prog
