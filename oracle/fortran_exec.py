"""oracle/fortran_exec.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Runs subroutines of the REFERENCE'S OWN Fortran source text (HYCOM-src: mod_tsadvc.F90, bigrid.F90, ...) without
a Fortran compiler: the structured subset of Fortran 90 those files use is translated, statement by statement,
into Python source and executed on numpy-backed arrays that keep the Fortran bounds and index order.  Nothing
about the algorithms is restated here - the translator only knows the language (do/if blocks, assignments,
expressions, calls) - so what it computes is what the reference text says, in IEEE double precision without
FMA contraction (Python floats), i.e. what a `-fdefault-real-8 -ffp-contract=off` build computes.

It exists to PIN the CPU oracle (oracle/tsadvc_oracle.c): tests/test_reference_text.py executes the reference
routines on small seeded cases and demands bit equality with the oracle; tests/golden/from_reference_text.json
keeps digests of those runs for machines where /root/reference is absent.

Supported: fixed/free form with `&` continuations, `!` comments, cpp (#if defined / #elif / #else / #endif /
#define NAME text), do / labelled do ... continue, block and one-line if, assignment (scalars, array elements,
whole arrays), call, return, parameter constants in declarations, intrinsics min max abs mod sign sqrt real int
nint float dble exp log atan2 cos sin, integer division, array-element actual arguments (sequence association:
the callee sees the array from that element on).  `if (c) go to L` forward to `L continue` at the same block level.  Not supported (raises): other goto, where, derived types, I/O
(write/print/read/open/close/flush are skipped), modules as such (module variables are entries of the
environment dictionary the caller supplies).
"""
from __future__ import annotations

import keyword
import math
import re

import numpy as np


# ---------------------------------------------------------------------------------------------------------
# arrays with Fortran bounds.  Storage is a numpy array in C order with the index order reversed - a(i,j,k) is
# store[k-lk, j-lj, i-li] - i.e. exactly the layout of the test arrays (nrows, ncols), which are wrapped, not copied
# ---------------------------------------------------------------------------------------------------------
class FArray:
    __slots__ = ("a", "lo", "rank", "isint")

    def __init__(self, store, lower):
        self.a = store
        self.lo = tuple(lower)
        self.rank = len(self.lo)
        assert store.ndim == self.rank, (store.shape, lower)
        self.isint = store.dtype.kind in "iub"

    @classmethod
    def zeros(cls, bounds, dtype=np.float64, fill=0):
        """bounds: ((lo, hi), ...) in Fortran order"""
        shape = tuple(hi - lo + 1 for lo, hi in reversed(bounds))
        return cls(np.full(shape, fill, dtype=dtype), [lo for lo, _ in bounds])

    def __getitem__(self, idx):
        if self.rank == 1:
            v = self.a[idx - self.lo[0]]
        elif self.rank == 2:
            v = self.a[idx[1] - self.lo[1], idx[0] - self.lo[0]]
        else:
            v = self.a[tuple(i - l for i, l in zip(reversed(idx), reversed(self.lo)))]
        return int(v) if self.isint else float(v)

    def __setitem__(self, idx, v):
        if self.rank == 1:
            self.a[idx - self.lo[0]] = v
        elif self.rank == 2:
            self.a[idx[1] - self.lo[1], idx[0] - self.lo[0]] = v
        else:
            self.a[tuple(i - l for i, l in zip(reversed(idx), reversed(self.lo)))] = v

    def fill(self, v):
        self.a[...] = v.a if isinstance(v, FArray) else v

    def section(self, idx):
        """numpy view of a(lo:hi, :, k, ...): `idx` holds ints (the dimension is dropped) or (lo, hi) pairs with
        None for an open end (the dimension stays); the view keeps the storage order (last Fortran dimension first)"""
        idx = idx if isinstance(idx, tuple) else (idx,)
        sel = []
        for d in range(self.rank - 1, -1, -1):
            x, l = idx[d], self.lo[d]
            if isinstance(x, tuple):
                lo, hi = x
                sel.append(slice(None if lo is None else lo - l, None if hi is None else hi - l + 1))
            else:
                sel.append(x - l)
        return self.a[tuple(sel)]

    def from_element(self, idx, rank):
        """sequence association: the array a callee with a `rank`-dimensional dummy sees when the actual argument is
        the element a(idx).  The first rank-1 dimensions keep their extent (the element must start them), everything
        from dimension `rank` on is one long last dimension that begins at the element - so that, as in Fortran,
        `call xctilr(saln(1-nbdy,1-nbdy,1,1),1,2*kk,..)` reaches both time levels.  The last dimension is numbered
        from the element's own index (callees that number it from 1 are told so by the harness)."""
        idx = tuple(idx) if isinstance(idx, tuple) else (idx,)
        lead = rank - 1
        if len(idx) == lead and self.rank == lead:      # a 2-D array handed to xctilr(a(1-nbdy,1-nbdy),1,1,..): one slab
            assert idx == self.lo
            return FArray(self.a[None], self.lo + (1,))
        assert all(i == l for i, l in zip(idx[:lead], self.lo[:lead])), "element is not the start of a slab"
        R = self.rank
        flat = self.a.reshape((-1,) + self.a.shape[R - lead:])
        tail = tuple(i - l for i, l in zip(reversed(idx[lead:]), reversed(self.lo[lead:])))
        start = int(np.ravel_multi_index(tail, self.a.shape[:R - lead]))
        return FArray(flat[start:], self.lo[:lead] + (idx[lead],))


# ---------------------------------------------------------------------------------------------------------
# source handling
# ---------------------------------------------------------------------------------------------------------
def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def load_source(path, defines=(), include_dirs=(), keep_omp=False):
    """logical, lower-cased statements of a Fortran file after cpp: list of (label or None, text).  keep_omp: the
    `!$OMP` directives come back as statements `$omp ...` (for the C backend, oracle/fortran_to_c.py)"""
    import os
    defined = {d: "" for d in defines}
    macros = {}
    raw = open(path, errors="replace").read().split("\n")
    out_lines, stack = [], []   # stack of [taking_now, any_branch_taken]

    def active():
        return all(s[0] for s in stack)
    i = 0
    while i < len(raw):
        ln = raw[i]
        i += 1
        s = ln.strip()
        if s.startswith("#"):
            d = re.sub(r"/\*.*?\*/", "", s[1:]).strip()
            m = re.match(r"(ifdef|ifndef|if|elif)\b\s*(.*)$", d)

            def cpp_true(e):
                # defined (X), (X), X joined by || && ! : an identifier is true when it is defined
                e = re.sub(r"defined\s*\(\s*(\w+)\s*\)", lambda q: " 1 " if (q.group(1) in defined or q.group(1) in macros) else " 0 ", e)
                e = re.sub(r"defined\s+(\w+)", lambda q: " 1 " if (q.group(1) in defined or q.group(1) in macros) else " 0 ", e)
                e = re.sub(r"[A-Za-z_]\w*", lambda q: " 1 " if (q.group(0) in defined or q.group(0) in macros) else " 0 ", e)
                e = e.replace("||", " or ").replace("&&", " and ").replace("!", " not ")
                return bool(eval(e, {"__builtins__": {}}))
            if m and m.group(1) in ("if", "ifdef", "ifndef"):
                t = cpp_true(m.group(2))
                t = (not t) if m.group(1) == "ifndef" else t
                stack.append([t, t])
            elif m and m.group(1) == "elif":
                t = (not stack[-1][1]) and cpp_true(m.group(2))
                stack[-1][0] = t
                stack[-1][1] = stack[-1][1] or t
            elif d.startswith("else"):
                stack[-1][0] = not stack[-1][1]
                stack[-1][1] = True
            elif d.startswith("endif"):
                stack.pop()
            elif d.startswith("define") and active():
                m = re.match(r"define\s+(\w+)\s*(.*)", d)
                macros[m.group(1)] = m.group(2).strip()
            elif d.startswith("include") and active():
                m = re.search(r'"([^"]+)"', d)
                for dd in (os.path.dirname(path),) + tuple(include_dirs):
                    f = os.path.join(dd, m.group(1))
                    if os.path.exists(f):
                        raw[i:i] = open(f, errors="replace").read().split("\n")
                        break
            continue
        if not active():
            continue
        mi = re.match(r"^\s*include\s+['\"]([^'\"]+)['\"]\s*(!.*)?$", ln, flags=re.I)
        if mi and not mi.group(1).lower().startswith("mpif"):       # the Fortran include line (stmt_fns.h in the shim)
            for dd in (os.path.dirname(path),) + tuple(include_dirs):
                f = os.path.join(dd, mi.group(1))
                if os.path.exists(f):
                    raw[i:i] = open(f, errors="replace").read().split("\n")
                    break
            continue
        out_lines.append(ln)
    # comments, continuations (free form: trailing &; the next line may start with &)
    stmts, cur = [], ""
    for ln in out_lines:
        omp = re.match(r"^\s*!(\$omp\b.*)$", ln, flags=re.I)
        if omp and keep_omp:
            ln = omp.group(1) if not cur else re.sub(r"^\$omp", "", omp.group(1), flags=re.I)
        else:
            ln = _strip_comment(ln)
        if not ln.strip():
            continue
        body = ln.strip()
        if cur:
            if body.startswith("&"):
                body = body[1:].lstrip()
            cur += " " + body
        else:
            cur = body
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        stmts.append(cur)
        cur = ""
    # macro substitution, lower case (strings are never needed), labels
    res = []
    for st in stmts:
        for _ in range(4):            # a macro may expand to another macro (BARRIER_MP -> BARRIER -> call ...)
            before = st
            for name, text in macros.items():
                st = re.sub(r"\b%s\b" % re.escape(name), text, st)
            if st == before:
                break
        st = re.sub(r"'[^']*'|\"[^\"]*\"", "''", st).lower()
        m = re.match(r"^(\d+)\s+(.*)$", st)
        lab, st = (int(m.group(1)), m.group(2)) if m else (None, st)
        parts = [p.strip() for p in st.split(";") if p.strip()]     # several statements on one line
        for k, p in enumerate(parts):
            res.append((lab if k == len(parts) - 1 else None, p))
    return res


def extract_unit(stmts, name):
    """(dummy argument names, body statements, internal procedures) of `subroutine name(...)`; the internal
    procedures (after `contains`) come back as a list of (name, args, body)"""
    name = name.lower()
    head = re.compile(r"^(?:recursive\s+)?subroutine\s+(\w+)\s*(?:\((.*)\))?\s*$")
    tail = re.compile(r"^end\s*(subroutine(\s+\w+)?)?\s*$")
    for k, (_, st) in enumerate(stmts):
        m = head.match(st)
        if m and m.group(1) == name:
            args = [a.strip() for a in (m.group(2) or "").split(",") if a.strip()]
            body, internals, cur, inside = [], [], None, False
            for lab, s2 in stmts[k + 1:]:
                if not inside:
                    if s2 == "contains":
                        inside = True
                        continue
                    if tail.match(s2):
                        return args, body, internals
                    body.append((lab, s2))
                else:
                    m2 = head.match(s2)
                    if cur is None and m2:
                        cur = (m2.group(1), [a.strip() for a in (m2.group(2) or "").split(",") if a.strip()], [])
                    elif cur is not None and tail.match(s2):
                        internals.append(cur)
                        cur = None
                    elif cur is not None:
                        cur[2].append((lab, s2))
                    elif tail.match(s2):
                        return args, body, internals
    raise KeyError(name)


# ---------------------------------------------------------------------------------------------------------
# expressions: tokenizer + precedence climbing -> Python source
# ---------------------------------------------------------------------------------------------------------
_TOK = re.compile(r"""\s*(?:
    (?P<str>'') |
    (?P<num>(?:\d+\.(?!(?:and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv|true|false)\.)\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?) |
    (?P<dotop>\.(?:and|or|not|eq|ne|lt|le|gt|ge|true|false|eqv|neqv)\.) |
    (?P<name>[a-z_]\w*(?:%[a-z_]\w*)*) |
    (?P<op>\*\*|==|/=|<=|>=|//|[-+*/(),<>:=])
)""", re.X)

_INTRINSIC = {
    "max": "_max", "min": "_min", "amax1": "_max", "amin1": "_min", "dmax1": "_max", "dmin1": "_min", "max0": "_max",
    "min0": "_min", "abs": "abs", "dabs": "abs", "iabs": "abs", "mod": "_mod", "sign": "_sign", "sqrt": "_sqrt",
    "dsqrt": "_sqrt", "real": "float", "float": "float", "dble": "float", "int": "_int", "nint": "_nint",
    "exp": "_exp", "log": "_log", "alog": "_log", "atan2": "_atan2", "cos": "_cos", "sin": "_sin", "atan": "_atan",
    "tan": "_tan", "acos": "_acos", "asin": "_asin", "minval": "_minval", "maxval": "_maxval", "merge": "_merge", "transfer": "_transfer",
}
_REL = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", "==": "==", "/=": "!=",
        "<": "<", "<=": "<=", ">": ">", ">=": ">="}


def _pyname(n):
    if "%" in n:                                   # a component of a derived-type variable: an attribute
        return ".".join(_pyname(x) for x in n.split("%"))
    return n + "_" if keyword.iskeyword(n) or n in ("print", "len", "id", "type", "sum", "all", "any") else n


def tokenize(s):
    pos, toks = 0, []
    s = s.strip()
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m or m.end() == pos:
            raise SyntaxError(f"cannot tokenize {s[pos:]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
    return toks


class ExprParser:
    """Fortran expression -> Python source.  `arrays`: names indexed with [] (everything else followed by '(' is a
    call); `rank`: rank of known arrays (for array-element actual arguments)"""

    def __init__(self, toks, arrays, funcs=(), callee=None):
        self.t, self.p, self.arrays, self.funcs, self.callee = toks, 0, arrays, funcs, callee or {}

    def peek(self):
        return self.t[self.p] if self.p < len(self.t) else (None, None)

    def take(self, val=None):
        k, v = self.peek()
        if val is not None and v != val:
            raise SyntaxError(f"expected {val!r}, got {v!r} in {self.t}")
        self.p += 1
        return k, v

    def parse(self):
        e = self.p_or()
        return e

    def p_or(self):
        e = self.p_and()
        while self.peek()[1] in (".or.",):
            self.take()
            e = f"({e} or {self.p_and()})"
        return e

    def p_and(self):
        e = self.p_not()
        while self.peek()[1] == ".and.":
            self.take()
            e = f"({e} and {self.p_not()})"
        return e

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.take()
            return f"(not {self.p_not()})"
        return self.p_rel()

    def p_rel(self):
        e = self.p_add()
        if self.peek()[1] in _REL:
            op = _REL[self.take()[1]]
            e = f"({e} {op} {self.p_add()})"
        return e

    def p_add(self):
        k, v = self.peek()
        if v in ("+", "-"):
            self.take()
            e = self.p_mul()
            e = f"(-{e})" if v == "-" else e
        else:
            e = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            e = f"({e} {op} {self.p_mul()})"
        return e

    def p_mul(self):
        e = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            r = self.p_pow()
            e = f"({e} * {r})" if op == "*" else f"_div({e}, {r})"
        return e

    def p_pow(self):
        b = self.p_atom()
        if self.peek()[1] == "**":
            self.take()
            k, v = self.peek()
            if v in ("+", "-"):
                self.take()
                ex = self.p_pow()
                ex = f"(-{ex})" if v == "-" else ex
            else:
                ex = self.p_pow()
            return f"_pow({b}, {ex})"
        return b

    def args(self):
        """comma list up to the closing parenthesis (consumed); a section lo:hi comes back as ('sec', lo, hi) with
        None for an open end"""
        out = []
        if self.peek()[1] == ")":
            self.take()
            return out
        while True:
            lo = None
            if self.peek()[1] != ":":
                lo = self.p_or()
            if self.peek()[1] == ":":
                self.take()
                hi = None
                if self.peek()[1] not in (",", ")"):
                    hi = self.p_or()
                out.append(("sec", lo, hi))
            else:
                out.append(lo)
            k, v = self.take()
            if v == ")":
                return out
            if v != ",":
                raise SyntaxError(f"expected , or ) got {v!r} in {self.t}")

    def p_atom(self):
        k, v = self.take()
        if k == "str":
            return "''"
        if k == "num":
            v = re.sub(r"_\w+$", "", v).replace("d", "e")
            return v if re.search(r"[.e]", v) else v.lstrip("0") or "0"
        if v == ".true.":
            return "True"
        if v == ".false.":
            return "False"
        if v == "(":
            e = self.p_or()
            self.take(")")
            return f"({e})"
        if k == "name":
            if self.peek()[1] == "(":
                self.take()
                a = self.args()
                if v in self.arrays or "%" in v:
                    if any(isinstance(x, tuple) for x in a):
                        parts = [f"({x[1]}, {x[2]})" if isinstance(x, tuple) else x for x in a]
                        return f"{_pyname(v)}.section(({', '.join(parts)},))"
                    return f"{_pyname(v)}[{', '.join(a)}]" if len(a) > 1 else f"{_pyname(v)}[{a[0]}]"
                if v in _INTRINSIC and v not in self.funcs:
                    return f"{_INTRINSIC[v]}({', '.join(a)})"
                if v in self.callee:       # a function with array dummies: an array-element actual is the array from there on
                    ranks = self.callee[v]
                    for k, x in enumerate(a):
                        me = re.match(r"^(\w+)\[(.*)\]$", x) if isinstance(x, str) else None
                        if me and k < len(ranks) and ranks[k]:
                            a[k] = f"{me.group(1)}.from_element(({me.group(2)},), {ranks[k]})"
                return f"{_pyname(v)}({', '.join(a)})"
            return _pyname(v)
        raise SyntaxError(f"unexpected {v!r} in {self.t}")


def expr(s, arrays, funcs=(), callee=None):
    p = ExprParser(tokenize(s), arrays, funcs, callee)
    e = p.parse()
    if p.p != len(p.t):
        raise SyntaxError(f"trailing tokens in {s!r}")
    return e


# ---------------------------------------------------------------------------------------------------------
# run-time helpers the generated code calls
# ---------------------------------------------------------------------------------------------------------
def _div(a, b):
    if isinstance(a, int) and isinstance(b, int) and not isinstance(a, bool):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    return a / b if b != 0 else (math.copysign(math.inf, a) * math.copysign(1.0, b) if a != 0 and a == a else math.nan)


def _pow(a, b):
    if isinstance(b, int):
        r = 1 if isinstance(a, int) else 1.0
        for _ in range(abs(b)):
            r = r * a
        return r if b >= 0 else _div(1.0, r)
    return math.pow(a, b)


def _max(*a):
    m = a[0]
    for x in a[1:]:
        if x > m:
            m = x
    return m


def _min(*a):
    m = a[0]
    for x in a[1:]:
        if x < m:
            m = x
    return m


def _mod(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return a - b * _div(a, b)
    return math.fmod(a, b)


def _sign(a, b):
    return abs(a) if (b > 0 or (b == 0 and math.copysign(1.0, b) > 0)) else -abs(a)


def _int(a):
    return int(a)


def _nint(a):
    return int(math.floor(a + 0.5)) if a >= 0 else -int(math.floor(-a + 0.5))


def _frange(a, b, c=1):
    return range(a, b + 1, c) if c > 0 else range(a, b - 1, c)


def _merge(t, f, mask):
    return t if mask else f


def _transfer(src, mold):
    """transfer(source, mold) between arrays: the bytes of `src` seen with the type of `mold`"""
    b = np.ascontiguousarray(src.a).view(np.uint8).reshape(-1)
    return FArray(b.view(mold.a.dtype)[:mold.a.size].copy(), mold.lo)


def _dummy(a, bounds):
    if not isinstance(a, FArray) or a.rank != len(bounds):
        return a
    lo = tuple(b[0] for b in bounds)
    n = bounds[-1][1] - bounds[-1][0] + 1
    if a.a.shape[0] > n > 0 or lo != a.lo:
        return FArray(a.a[:n] if a.a.shape[0] > n > 0 else a.a, lo)
    return a


class FortranStop(Exception):
    pass


RUNTIME = dict(np=np, _dummy=_dummy, _merge=_merge, _transfer=_transfer, _minval=lambda a: float(np.min(a)), _maxval=lambda a: float(np.max(a)), _div=_div, _pow=_pow, _max=_max, _min=_min, _mod=_mod, _sign=_sign, _int=_int, _nint=_nint,
               _frange=_frange, _sqrt=math.sqrt, _exp=math.exp, _log=math.log, _atan2=math.atan2, _cos=math.cos,
               _sin=math.sin, _atan=math.atan, _tan=math.tan, _acos=math.acos, _asin=math.asin, FArray=FArray,
               FortranStop=FortranStop)

_DECL = re.compile(r"^(real|integer|logical|character|double\s*precision|implicit|use|save|external|intrinsic|"
                   r"dimension|private|public|data|common|parameter|type|include|allocatable|intent)\b")
_IO = re.compile(r"^(write|print|read|open|close|flush|rewind|format|call\s+flush)\b")


def _split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return [x.strip() for x in out]


def _match_paren(s, start):
    depth = 0
    for k in range(start, len(s)):
        if s[k] == "(":
            depth += 1
        elif s[k] == ")":
            depth -= 1
            if depth == 0:
                return k
    raise SyntaxError(s)


class Translator:
    """one subroutine -> Python source.  `env_arrays`: rank of every array visible to the unit (dummies included);
    `module_scalars`: module variables the unit may assign (declared global); `callee_ranks`: for a subroutine
    called from here, the ranks of its dummy arrays by position (None for scalars) so that array-element actual
    arguments become views; `skip_calls`: calls to drop; `inout_calls`: `call f(x)` that means x = f(x)"""

    def __init__(self, name, args, body, env_arrays, module_scalars=(), callee_ranks=None, skip_calls=(),
                 inout_calls=(), funcs=(), drop_blocks=(), internals=(), host=None):
        self.name, self.args, self.body = name, args, body
        self.arr = dict(env_arrays)
        self.modsc = set(module_scalars)
        self.callee = callee_ranks or {}
        self.skip = set(skip_calls)
        self.inout = set(inout_calls)
        self.funcs = set(funcs)
        self.drop = [re.compile(r) for r in drop_blocks]   # block-ifs (diagnostic output) to leave out entirely
        self.internals, self.host = list(internals), host
        self.kw = dict(callee_ranks=callee_ranks, skip_calls=skip_calls, inout_calls=inout_calls, drop_blocks=drop_blocks)
        self.lines, self.ind = [], 1
        self.assigned = set()
        self.do_labels = []   # stack of labels of open labelled do loops (None for unlabelled)
        self.declared = set()  # scalars the unit declares itself (locals of an internal procedure, not the host's)
        self.goto_labels = []  # stack of (label, indentation) of open forward jumps

    def emit(self, s):
        self.lines.append("    " * self.ind + s)

    def ex(self, s):
        return expr(s, self.arr, self.funcs, self.callee)

    def stmt(self, st):
        m = re.match(r"^if\s*\((.*)\)\s*go\s*to\s*(\d+)$", st)
        if m:
            # a forward jump over the statements up to `label continue` at the same block level (cnuity.F90:749, 978):
            # they become the body of `if not (condition)`
            self.emit(f"if not {self.ex(m.group(1))}:")
            self.ind += 1
            self.emit("pass")
            self.goto_labels.append((int(m.group(2)), self.ind))
            return
        if re.match(r"^(go\s*to|where|forall|select|cycle|exit)\b", st):
            raise NotImplementedError(st)
        if _IO.match(st) or st in ("continue",):
            self.emit("pass")
            return
        m = re.match(r"^type\s*\(\s*(\w+)\s*\)\s*(?:,[^:]*)?::(.*)$", st)
        if m and m.group(1) != "c_ptr":               # a variable of a derived type: the harness builds it (_newtype)
            for n in _split_top(m.group(2)):
                self.emit(f"{_pyname(n.strip())} = _newtype('{m.group(1)}')")
            return
        m = re.match(r"^character\b.*::\s*(\w+)\s*\((\w+)\)$", st)
        if m:                                          # character(kind=c_char) :: id(128): bytes
            self.arr[m.group(1)] = 1
            self.emit(f"{_pyname(m.group(1))} = FArray.zeros(((1, {self.ex(m.group(2))}),), dtype=np.uint8)")
            return
        m = re.match(r"^allocate\s*\((.*)\)$", st)
        if m:
            for ent in _split_top(m.group(1)):
                me = re.match(r"^(\w+)\s*\((.*)\)$", ent)
                bb = []
                for b in _split_top(me.group(2)):
                    lo, hi = (b.split(":") + [None])[:2] if ":" in b else ("1", b)
                    bb.append(f"({self.ex(lo)}, {self.ex(hi)})")
                self.assigned.add(me.group(1))
                self.emit(f"{_pyname(me.group(1))} = FArray.zeros(({', '.join(bb)},))")
            return
        m = re.match(r"^(real|integer|logical|double\s*precision)\b(.*?)::(.*)$", st)
        if ((m and "parameter" in m.group(2)) or re.match(r"^parameter\s*\(", st)) and getattr(self, "params_done", False):
            return            # emitted ahead of the local arrays by declarations()
        if m and "parameter" in m.group(2):
            for item in _split_top(m.group(3)):
                if "=" in item:
                    n, v = item.split("=", 1)
                    if "(/" in v:       # a constant vector: real, parameter, dimension(7) :: c = (/ .. /)
                        vals = _split_top(v[v.index("(/") + 2:v.rindex("/)")])
                        self.arr[n.strip()] = 1
                        self.emit(f"{_pyname(n.strip())} = FArray(np.array([{', '.join(self.ex(x) for x in vals)}], dtype=np.float64), (1,))")
                        continue
                    self.emit(f"{_pyname(n.strip())} = {self.ex(v)}")
            return
        m = re.match(r"^parameter\s*\((.*)\)$", st)
        if m:
            for item in _split_top(m.group(1)):
                n, v = item.split("=", 1)
                self.emit(f"{_pyname(n.strip())} = {self.ex(v)}")
            return
        if _DECL.match(st):
            return
        if st == "return":
            self.emit("return")
            return
        if st.startswith("stop"):
            self.emit("raise FortranStop()")
            return
        m = re.match(r"^do\s+(?:(\d+)\s+)?(\w+)\s*=\s*(.*)$", st)
        if m:
            parts = _split_top(m.group(3))
            rng = ", ".join(self.ex(p) for p in parts)
            self.nloops = getattr(self, "nloops", 0) + 1
            rname, var = f"_r{self.nloops}", _pyname(m.group(2))
            self.emit(f"{rname} = _frange({rng})")
            self.emit(f"for {var} in {rname}:")
            self.ind += 1
            self.emit("pass")
            self.assigned.add(m.group(2))
            self.do_labels.append((int(m.group(1)) if m.group(1) else None, var, rname))
            return
        if re.match(r"^end\s*do$", st):
            self.close_do()
            return
        m = re.match(r"^(else\s*if|elseif|if)\s*\(", st)
        if m and m.group(1) == "if" and "if" in self.arr:
            e0 = _match_paren(st, st.index("("))
            if re.match(r"^\s*=[^=]", st[e0 + 1:]):
                m = None          # an array called `if` (bigrid.F90: indxi) is being assigned
        if m:
            start = st.index("(", m.end() - 1)
            end = _match_paren(st, start)
            cond, rest = st[start + 1:end], st[end + 1:].strip()
            kw = "if" if m.group(1) == "if" else "elif"
            if rest == "then":
                if kw == "elif":
                    self.ind -= 1
                self.emit(f"{kw} {self.ex(cond)}:")
                self.ind += 1
                self.emit("pass")
            else:
                assert kw == "if", st
                self.emit(f"if {self.ex(cond)}:")
                self.ind += 1
                self.stmt(rest)
                self.ind -= 1
            return
        if st == "else":
            self.ind -= 1
            self.emit("else:")
            self.ind += 1
            self.emit("pass")
            return
        if re.match(r"^end\s*if$", st):
            self.ind -= 1
            return
        m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", st)
        if m:
            f, a = m.group(1), (_split_top(m.group(2)) if m.group(2) else [])
            if f in self.skip:
                self.emit("pass")
                return
            if f in self.inout:
                self.emit(f"{_pyname(a[0])} = {f}({', '.join(self.ex(x) for x in a)})")
                self.assigned.add(a[0])
                return
            ranks = self.callee.get(f)
            out = []
            for k, x in enumerate(a):
                mm = re.match(r"^(\w+)\s*\((.*)\)$", x)
                if mm and mm.group(1) in self.arr and ranks is not None and k < len(ranks) and ranks[k]:
                    idx = ", ".join(self.ex(y) for y in _split_top(mm.group(2)))
                    out.append(f"{_pyname(mm.group(1))}.from_element(({idx},), {ranks[k]})")
                else:
                    out.append(self.ex(x))
            self.emit(f"{_pyname(f)}({', '.join(out)})")
            return
        # assignment
        depth, eq = 0, -1
        for k, ch in enumerate(st):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and st[k - 1] not in "<>/=" and st[k + 1:k + 2] != "=":
                eq = k
                break
        if eq < 0:
            raise NotImplementedError(st)
        lhs, rhs = st[:eq].strip(), st[eq + 1:].strip()
        m = re.match(r"^([\w%]+)\s*\((.*)\)$", lhs)
        if m:
            n, idx = m.group(1), _split_top(m.group(2))
            if n not in self.arr and "%" not in n:
                if all(re.match(r"^[a-z_]\w*$", i) for i in idx):   # a statement function: name(dummies) = expression
                    self.funcs.add(n)
                    self.emit(f"def {_pyname(n)}({', '.join(_pyname(i) for i in idx)}):")
                    self.emit(f"    return {self.ex(rhs)}")
                    return
                raise NotImplementedError(f"assignment to unknown array {n}: {st}")
            if all(i == ":" for i in idx):
                self.emit(f"{_pyname(n)}.fill({self.ex(rhs)})")
            elif any(":" in i for i in idx):
                self.emit(f"{self.ex(lhs)}[...] = {self.ex(rhs)}")
            else:
                ii = ", ".join(self.ex(i) for i in idx)
                self.emit(f"{_pyname(n)}[{ii}] = {self.ex(rhs)}")
        elif lhs in self.arr:
            self.emit(f"{_pyname(lhs)}.fill({self.ex(rhs)})")
        else:
            self.assigned.add(lhs)
            self.emit(f"{_pyname(lhs)} = {self.ex(rhs)}")

    def close_do(self):
        """end of a do loop: the variable is left one step past the last iteration, as in Fortran"""
        _, var, rname = self.do_labels.pop()
        self.ind -= 1
        self.emit(f"{var} = {rname}.start + len({rname}) * {rname}.step")

    def declarations(self):
        """arrays declared in the unit: dummies get their rank; locals with explicit bounds are allocated (after the
        parameter constants, which may size them)"""
        # an undefined local scalar may be read in Fortran (cnuity.F90:1006 reads iflip, which is only set when thkdf4
        # is used, to fill an array nobody reads): NaN / a sentinel instead of a Python error.  Ahead of the parameter
        # constants, which overwrite the entry of a name that is typed first and given its value later.
        saved = set()          # `save a, b` / `integer, save :: a`: such a variable lives in the environment the caller
        for _, st in self.body:    # supplies (with the value its `data` statement gives it) and is never reset here
            m = re.match(r"^save\s+(.*)$", st)
            if m:
                saved |= {x.strip() for x in m.group(1).split(",")}
            m = re.match(r"^(real|integer|logical|double\s*precision)\b(.*?)::(.*)$", st)
            if m and re.search(r"\bsave\b", m.group(2)):
                saved |= {re.sub(r"\(.*", "", x).strip() for x in _split_top(m.group(3))}
        for _, st in self.body:
            m = re.match(r"^(real|integer|logical|double\s*precision)\b(.*)$", st)
            if not m or "parameter" in m.group(2).split("::")[0] or "dimension" in m.group(2).split("::")[0]:
                continue
            undef = "-987654321" if m.group(1) == "integer" else ("False" if m.group(1) == "logical" else "float('nan')")
            for ent in _split_top(m.group(2).split("::", 1)[-1]):
                self.declared.add(ent)
                if re.match(r"^[a-z_]\w*$", ent) and ent not in self.args and ent not in saved and \
                        (ent not in self.arr or self.host is None):
                    self.emit(f"{_pyname(ent)} = {undef}")
        self.params_done = True
        for _, st in self.body:
            m = re.match(r"^(real|integer|logical|double\s*precision)\b(.*?)::(.*)$", st)
            if (m and "parameter" in m.group(2)) or re.match(r"^parameter\s*\(", st):
                self.params_done = False
                self.stmt(st)
        self.params_done = True
        for _, st in self.body:
            m = re.match(r"^(real|integer|logical|double\s*precision)\b(.*)$", st)
            if not m or "parameter" in m.group(2).split("::")[0]:
                continue
            isint = m.group(1) in ("integer", "logical")
            rest = m.group(2)
            if "::" in rest:
                attrs, ents = rest.split("::", 1)
            else:
                attrs, ents = "", rest
            dim = None
            md = re.search(r"dimension\s*\(", attrs)
            if md:
                end = _match_paren(attrs, md.end() - 1)
                dim = _split_top(attrs[md.end():end])
            for ent in _split_top(ents):
                me = re.match(r"^(\w+)\s*(?:\((.*)\))?(?:\s*\*\s*\d+)?$", ent.split("=")[0].strip())
                if not me:
                    continue
                n, own = me.group(1), me.group(2)
                bounds = _split_top(own) if own else dim
                if not bounds:
                    if n in self.arr and n not in self.args and self.host is None:
                        del self.arr[n]       # a local scalar that hides a module array of the same name
                    continue
                known = n in self.arr
                self.arr[n] = len(bounds)
                explicit = all(":" not in b or b.count(":") == 1 for b in bounds) and not any(b.strip() in (":", "*") for b in bounds)
                if n in self.args and explicit and self.host is None:
                    # a dummy array has the bounds ITS declaration gives it: lower bounds, and the extent of the last
                    # dimension (the actual argument may be a longer sequence-associated view)
                    bb = []
                    for b in bounds:
                        lo, hi = (b.split(":") + [None])[:2] if ":" in b else ("1", b)
                        bb.append(f"({self.ex(lo)}, {self.ex(hi)})")
                    self.emit(f"{_pyname(n)} = _dummy({_pyname(n)}, ({', '.join(bb)},))")
                if not known and n not in self.args and "allocatable" not in attrs and all(":" not in b or b.count(":") == 1 for b in bounds) \
                        and not any(b.strip() in (":", "*") for b in bounds):
                    bb = []
                    for b in bounds:
                        lo, hi = (b.split(":") + [None])[:2] if ":" in b else ("1", b)
                        bb.append(f"({self.ex(lo)}, {self.ex(hi)})")
                    self.emit(f"{_pyname(n)} = FArray.zeros(({', '.join(bb)},), dtype={'np.int64' if isint else 'np.float64'})")

    def source(self):
        self.declarations()
        skipping = 0
        for lab, st in self.body:
            if skipping:
                if re.match(r"^if\s*\(.*\)\s*then$", st):
                    skipping += 1
                elif re.match(r"^end\s*if$", st):
                    skipping -= 1
                continue
            if any(r.search(st) for r in self.drop) and re.match(r"^if\s*\(.*\)\s*then$", st):
                skipping = 1
                continue
            self.stmt(st)
            while lab is not None and self.do_labels and self.do_labels[-1][0] == lab:
                self.close_do()
            if lab is not None and self.goto_labels and self.goto_labels[-1][0] == lab:
                assert self.goto_labels[-1][1] == self.ind, "a jump into or out of a block"
                self.ind -= 1
                self.goto_labels.pop()
        if self.host is not None:
            return None
        # internal procedures (host association): nested functions that see the host's variables; what they assign
        # and the host assigns too is the host's variable
        inner_src = []
        for iname, iargs, ibody in self.internals:
            t = Translator(iname, iargs, ibody, self.arr, module_scalars=self.modsc, funcs=self.funcs, host=self, **self.kw)
            t.ind = 2
            t.source()
            shared = sorted(_pyname(n) for n in t.assigned if (n in self.assigned or n in self.args or n in self.declared) and n not in iargs and n not in t.declared)
            inner_src.append(f"    def {_pyname(iname)}({', '.join(_pyname(a) for a in iargs)}):")
            if shared:
                inner_src.append("        nonlocal " + ", ".join(shared))
            inner_src.append("        pass")
            inner_src += t.lines
            self.assigned |= {n for n in t.assigned if n in self.modsc}
        glob = sorted(_pyname(n) for n in self.assigned if n in self.modsc and n not in self.args)
        head = [f"def {self.name}({', '.join(_pyname(a) for a in self.args)}):"]
        if glob:
            head.append("    global " + ", ".join(glob))
        # statement functions and parameters come first in self.lines; internal procedures may use them, so they go
        # after the declarations block: Python resolves the names when the internal procedure is called
        return "\n".join(head + ["    pass"] + inner_src + self.lines) + "\n"


def compile_slice(path, unit, first, last, name, env, defines=("RELO",), back=0, **kw):
    """a run of statements of `subroutine unit` as a parameterless subroutine `name`: from the first statement that
    matches the regular expression `first` up to (not including) the first later statement that matches `last`.
    For routines whose head cannot be executed (xcspmd: file input, MPI start-up) but whose tail is plain Fortran."""
    stmts = load_source(path, defines)
    _, body, _ = extract_unit(stmts, unit)
    k0 = next(k for k, (_, st) in enumerate(body) if re.search(first, st)) - back     # `back`: enclosing do statements
    k1 = next(k for k, (_, st) in enumerate(body) if k > k0 and re.search(last, st))
    arrays = {k: v.rank for k, v in env.items() if isinstance(v, FArray)}
    t = Translator(name.lower(), [], body[k0:k1], arrays, module_scalars=[k for k, v in env.items() if not isinstance(v, FArray)], **kw)
    src = t.source()
    for k, v in RUNTIME.items():
        env.setdefault(k, v)
    exec(compile(src, f"<{name}: part of {unit} of {path}>", "exec"), env)
    return src


def compile_unit(path, name, env, defines=("RELO",), extra_arrays=None, include_dirs=(), **kw):
    """translate `subroutine name` of the file and define it in `env` (a dict that holds the module variables:
    FArray objects and scalars); returns the Python source (for inspection)"""
    stmts = load_source(path, defines, include_dirs)
    args, body, internals = extract_unit(stmts, name)
    arrays = {k: v.rank for k, v in env.items() if isinstance(v, FArray)}
    arrays.update(extra_arrays or {})
    t = Translator(name.lower(), args, body, arrays, module_scalars=[k for k, v in env.items() if not isinstance(v, FArray)],
                   internals=internals, **kw)
    src = t.source()
    for k, v in RUNTIME.items():
        env.setdefault(k, v)
    exec(compile(src, f"<{name} of {path}>", "exec"), env)
    return src
