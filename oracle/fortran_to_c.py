"""oracle/fortran_to_c.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The second backend of the Fortran-subset translator (oracle/fortran_exec.py is the first): the same statements of
the REFERENCE'S OWN SOURCE TEXT, read from /root/reference where they lie, are turned into C (gcc: nested functions
for statement functions and internal procedures, statement expressions for the intrinsics) and compiled into
oracle/_ref/libref_text.so.  Nothing of the algorithms is restated: the translator knows the language only.  What it
is for:
  * the pin at full size - the interpreter needs a minute for the 150 x 150 x 22 box, the compiled text runs a
    4500 x 3298 layer in seconds (tests/test_reference_text_c.py: compiled text == interpreted text == oracle);
  * `bench.py --impl reference`: the reference's own sweeps, `!$OMP PARALLEL DO ... SCHEDULE(STATIC,jblk)` carried
    over as `#pragma omp parallel for ... schedule(static,jblk)`, timed on the host cores.
Arithmetic: IEEE double, gcc -O2 -ffp-contract=off for the parity build (what `gfortran -O2 -ffp-contract=off` does to
the same statements), max/min keep the first argument on a tie like gfortran's inline expansion, x**n with integer n
multiplies left to right like the interpreter.

Every Fortran name becomes f_<name> (j0, y1, index ... are libm / libc names).  Module variables are C globals set
from the interpreter's environment (bind()): a scalar is `int` / `double`, an array `double *` / `int32_t *` with its
lower bounds and extents in f_<name>_lo[] / f_<name>_n[].
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

import fortran_exec as fx
from fortran_exec import FArray, _IO, _match_paren, _split_top, tokenize

PRELUDE = r"""
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define MAX2_(a,b) ({ __typeof__((a)+(b)) _a = (a), _b = (b); _b > _a ? _b : _a; })
#define MIN2_(a,b) ({ __typeof__((a)+(b)) _a = (a), _b = (b); _b < _a ? _b : _a; })
static inline long long iabs_(long long a) { return a < 0 ? -a : a; }
#define ABS_(a) _Generic((a), double: fabs, float: fabsf, default: iabs_)(a)
static inline long long imod_(long long a, long long b) { return a % b; }
#define MOD_(a,b) _Generic((a)+(b), double: fmod, default: imod_)((a),(b))
static inline double dsign_(double a, double b) { return copysign(fabs(a), b); }
static inline long long isign_(long long a, long long b) { return b >= 0 ? iabs_(a) : -iabs_(a); }
#define SIGN_(a,b) _Generic((a)+(b), double: dsign_, default: isign_)((a),(b))
static inline double dpow_i(double a, int n) { double r = 1.0; for (int k = 0; k < (n < 0 ? -n : n); k++) r = r * a; return n >= 0 ? r : 1.0 / r; }
static inline long long ipow_i(long long a, int n) { long long r = 1; for (int k = 0; k < n; k++) r = r * a; return r; }
#define POW_(a,b) _Generic((b), int: _Generic((a), double: dpow_i, default: ipow_i), default: pow)((a),(b))
#define NINT_(a) ((int)lround(a))
static void free_(void *p) { free(*(void **)p); }
static int f_stop_count_ = 0;
static void f_stop_(void) { f_stop_count_++; }
int ref_text_stop_count(void) { return f_stop_count_; }
"""

_CT = {"real": "double", "doubleprecision": "double", "integer": "int", "logical": "int"}
_INTR = {"max": "MAX", "min": "MIN", "amax1": "MAX", "amin1": "MIN", "dmax1": "MAX", "dmin1": "MIN", "max0": "MAX",
         "min0": "MIN", "abs": "ABS_", "dabs": "ABS_", "iabs": "ABS_", "mod": "MOD_", "sign": "SIGN_", "sqrt": "sqrt",
         "dsqrt": "sqrt", "real": "(double)", "float": "(double)", "dble": "(double)", "int": "(int)", "nint": "NINT_",
         "exp": "exp", "log": "log", "alog": "log", "atan2": "atan2", "cos": "cos", "sin": "sin", "atan": "atan",
         "tan": "tan", "acos": "acos", "asin": "asin"}
_REL = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", "==": "==", "/=": "!=",
        "<": "<", "<=": "<=", ">": ">", ">=": ">="}


class Arr:
    """an array in scope: how to index it (`lo`, `n`: C expressions per dimension; n of the last may be None)"""

    def __init__(self, cname, ctype, lo, n):
        self.cname, self.ctype, self.lo, self.n, self.rank = cname, ctype, lo, n, len(lo)

    def elem(self, idx):
        assert len(idx) == self.rank, (self.cname, idx)
        e = f"(({idx[-1]})-({self.lo[-1]}))"
        for d in range(self.rank - 2, -1, -1):
            e = f"((({idx[d]})-({self.lo[d]})) + ({self.n[d]})*{e})"
        return f"{self.cname}[{e}]"


class CExpr:
    """Fortran expression -> C; `scope`: name -> Arr; `funcs`: names that are functions of the unit (statement
    functions); `sect`: loop variables for the sectioned dimensions of an array assignment"""

    def __init__(self, toks, scope, funcs, sect=None):
        self.t, self.p, self.scope, self.funcs, self.sect = toks, 0, scope, funcs, sect

    def peek(self):
        return self.t[self.p] if self.p < len(self.t) else (None, None)

    def take(self, val=None):
        k, v = self.peek()
        if val is not None and v != val:
            raise SyntaxError(f"expected {val!r}, got {v!r} in {self.t}")
        self.p += 1
        return k, v

    def p_or(self):
        e = self.p_and()
        while self.peek()[1] == ".or.":
            self.take()
            e = f"({e} || {self.p_and()})"
        return e

    def p_and(self):
        e = self.p_not()
        while self.peek()[1] == ".and.":
            self.take()
            e = f"({e} && {self.p_not()})"
        return e

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.take()
            return f"(!{self.p_not()})"
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
            e = f"({e} {op} {self.p_pow()})"
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
            return f"POW_({b}, {ex})"
        return b

    def args(self):
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

    def section_elem(self, arr, a):
        """element of `arr` inside the loops of an array assignment: its k-th sectioned dimension runs with the
        k-th loop variable (an offset from 0)"""
        if self.sect is None:
            raise NotImplementedError(f"array section of {arr.cname} outside an array assignment")
        idx, k = [], 0
        for d, x in enumerate(a):
            if isinstance(x, tuple):
                lo = x[1] if x[1] is not None else arr.lo[d]
                idx.append(f"(({lo}) + {self.sect[k]})")
                k += 1
            else:
                idx.append(x)
        return arr.elem(idx)

    def p_atom(self):
        k, v = self.take()
        if k == "str":
            return "0"
        if k == "num":
            v = re.sub(r"_\w+$", "", v).replace("d", "e")
            if re.search(r"[.e]", v):
                return v
            return v.lstrip("0") or "0"
        if v == ".true.":
            return "1"
        if v == ".false.":
            return "0"
        if v == "(":
            e = self.p_or()
            self.take(")")
            return f"({e})"
        if k == "name":
            if self.peek()[1] == "(":
                self.take()
                a = self.args()
                if v in self.scope:
                    arr = self.scope[v]
                    if any(isinstance(x, tuple) for x in a):
                        if self.sect is None and arr.rank == 1 and a[0][1] is not None and a[0][2] is not None:
                            # x(lo:hi) as the argument of minval / maxval: a reduction loop
                            lo, hi = a[0][1], a[0][2]
                            return (f"@sec:({{ {arr.ctype} _m = {arr.elem([lo])}; for (int _q = ({lo}) + 1; _q <= ({hi}); _q++) "
                                    f"_m = @OP(_m, {arr.elem(['_q'])}); _m; }})")
                        return self.section_elem(arr, a)
                    return arr.elem(a)
                if v in ("minval", "maxval") and len(a) == 1 and isinstance(a[0], str) and a[0].startswith("@sec:"):
                    return a[0][5:].replace("@OP", "MIN2_" if v == "minval" else "MAX2_")
                if v in _INTR and v not in self.funcs:
                    f = _INTR[v]
                    if f in ("MAX", "MIN"):
                        e = a[0]
                        for x in a[1:]:
                            e = f"{f}2_({e}, {x})"
                        return e
                    if f.startswith("("):
                        return f"({f}({a[0]}))"
                    return f"{f}({', '.join(a)})"
                return f"f_{v}({', '.join(a)})"
            if v in self.scope and self.sect is not None:       # a whole array inside an array assignment
                arr = self.scope[v]
                return self.section_elem(arr, [("sec", None, None)] * arr.rank)
            return f"f_{v}"
        raise SyntaxError(f"unexpected {v!r} in {self.t}")


def _parse_decl(st):
    """(type, attributes, [(name, bounds or None, initial value or None)]) of a type declaration, else None"""
    m = re.match(r"^(real|integer|logical|double\s*precision)\b(.*)$", st)
    if not m:
        return None
    ty, rest = m.group(1).replace(" ", ""), m.group(2)
    rest = re.sub(r"^\s*\*\s*\d+", "", rest)
    attrs, ents = rest.split("::", 1) if "::" in rest else ("", rest)
    dim = None
    md = re.search(r"dimension\s*\(", attrs)
    if md:
        end = _match_paren(attrs, md.end() - 1)
        dim = _split_top(attrs[md.end():end])
    out = []
    for ent in _split_top(ents):
        val = None
        if "=" in ent and "(/" in ent:
            ent, val = ent.split("=", 1)
        elif "=" in ent:
            ent, val = ent.split("=", 1)
        me = re.match(r"^(\w+)\s*(?:\((.*)\))?(?:\s*\*\s*\d+)?$", ent.strip())
        if not me:
            continue
        out.append((me.group(1), _split_top(me.group(2)) if me.group(2) else dim, val.strip() if val else None))
    return ty, attrs, out


class CUnit:
    """one subroutine (or internal procedure) -> C function text"""

    def __init__(self, gen, name, args, body, internals=(), host=None):
        self.gen, self.name, self.args, self.body, self.internals, self.host = gen, name, args, body, internals, host
        self.scope = dict(host.scope if host else gen.modarr)       # arrays in scope
        self.funcs = set(host.funcs if host else ())
        self.lines, self.ind = [], 1
        self.do_stack, self.nloop = [], 0
        self.decl = {}          # name -> (ctype, bounds, value, attrs)
        for _, st in body:
            d = _parse_decl(st)
            if d:
                for n, b, v in d[2]:
                    self.decl[n] = (_CT[d[0]], b, v, d[1])
            m = re.match(r"^parameter\s*\((.*)\)$", st)
            if m:
                for item in _split_top(m.group(1)):
                    n, v = item.split("=", 1)
                    n = n.strip()
                    ct, b, _, at = self.decl.get(n, ("double", None, None, ""))
                    self.decl[n] = (ct, b, v.strip(), at + " parameter")
        # statement functions: name(dummies) = expression, where name is no array
        self.stfuncs = set()
        for _, st in body:
            m = re.match(r"^(\w+)\s*\(([\w\s,]*)\)\s*=(?!=)", st)
            if m and m.group(1) not in self.scope and not (m.group(1) in self.decl and self.decl[m.group(1)][1] is not None):
                self.stfuncs.add(m.group(1))

    # -- helpers -------------------------------------------------------------------------------------------
    def emit(self, s):
        self.lines.append("  " * self.ind + s)

    def ex(self, s, sect=None):
        p = CExpr(tokenize(s), self.scope, self.funcs, sect)
        e = p.p_or()
        if p.p != len(p.t):
            raise SyntaxError(f"trailing tokens in {s!r}")
        return e

    def is_array_dummy(self, n):
        return n in self.decl and self.decl[n][1] is not None

    def bounds_c(self, bounds):
        lo, nn = [], []
        for b in bounds:
            b = b.strip()
            if b in ("*", ":"):
                lo.append("1")
                nn.append(None)
                continue
            l, h = (b.split(":") + [None])[:2] if ":" in b else ("1", b)
            l = l.strip() or "1"
            lo.append(self.ex(l))
            nn.append(None if not h or h.strip() == "*" else f"(({self.ex(h)})-({self.ex(l)})+1)")
        return lo, nn

    # -- declarations ---------------------------------------------------------------------------------------
    def header(self):
        ps = []
        for a in self.args:
            ct, b, _, _ = self.decl.get(a, ("double", None, None, ""))
            if b is not None:
                ps.append(f"{'int32_t' if ct == 'int' else 'double'} *f_{a}")
            else:
                ps.append(f"{ct} f_{a}")
        return f"void f_{self.name}({', '.join(ps) or 'void'})"

    def declarations(self):
        # dummies first (their bounds may only use module variables and other dummies), then parameters and locals in
        # the order of the text
        for a in self.args:
            if self.is_array_dummy(a):
                ct, b, _, _ = self.decl[a]
                lo, nn = self.bounds_c(b)
                self.scope[a] = Arr(f"f_{a}", "int32_t" if ct == "int" else "double", lo, nn)
        done = set(self.args)
        for _, st in self.body:
            names = []
            d = _parse_decl(st)
            if d:
                names = [n for n, _, _ in d[2]]
            m = re.match(r"^parameter\s*\((.*)\)$", st)
            if m:
                names = [item.split("=", 1)[0].strip() for item in _split_top(m.group(1))]
            for n in names:
                ct, b, v, at = self.decl[n]
                is_param = "parameter" in at
                if n in done or n in self.stfuncs:
                    continue
                done.add(n)
                if "allocatable" in at:
                    if n not in self.gen.modarr:
                        raise KeyError(f"{self.name}: allocatable {n} must be supplied as a module array")
                    continue
                if b is None:
                    if n in self.scope and self.host is None:
                        del self.scope[n]         # a local scalar hides a module array of the same name
                    if is_param:
                        self.emit(f"const {ct} f_{n} = {self.ex(v)};")
                    elif "save" in at or n in self.saved:
                        if n not in self.gen.modsc:
                            raise KeyError(f"{self.name}: saved {n} must be supplied as a module scalar")
                    else:
                        undef = "-987654321" if ct == "int" else "NAN"
                        self.emit(f"{ct} f_{n} = {'0' if d and d[0] == 'logical' else undef}; (void)f_{n};")
                elif is_param and v is not None and "(/" in v:
                    vals = _split_top(v[v.index("(/") + 2:v.rindex("/)")])
                    self.emit(f"const double f_{n}[] = {{{', '.join(self.ex(x) for x in vals)}}};")
                    self.scope[n] = Arr(f"f_{n}", "double", ["1"], [None])
                elif n in self.gen.modarr and self.host is None:
                    continue                      # a local (or saved) array the environment supplies, to be looked at
                                                  # after the call (xmin, xmax of tsadvc; like the interpreter)
                else:
                    lo, nn = self.bounds_c(b)
                    at_ = "int32_t" if ct == "int" else "double"
                    size = " * ".join(x for x in nn)
                    self.emit(f"{at_} *f_{n} __attribute__((cleanup(free_))) = calloc(({size}) > 0 ? ({size}) : 1, sizeof({at_}));")
                    self.scope[n] = Arr(f"f_{n}", at_, lo, nn)

    # -- statements -----------------------------------------------------------------------------------------
    def array_assign(self, lhs_name, lhs_idx, rhs):
        arr = self.scope[lhs_name]
        idx = lhs_idx if lhs_idx is not None else [":"] * arr.rank
        loops, lidx = [], []
        for d, x in enumerate(idx):
            x = x.strip()
            if ":" in x:
                l, h = [y.strip() for y in x.split(":")]
                lo = self.ex(l) if l else arr.lo[d]
                cnt = f"(({self.ex(h)})-({lo})+1)" if h else (f"(({arr.n[d]})-(({lo})-({arr.lo[d]})))" if arr.n[d] else None)
                if cnt is None:
                    raise NotImplementedError(f"section of an assumed-size dimension of {lhs_name}")
                v = f"_s{len(loops) + 1}"
                loops.append((v, cnt))
                lidx.append(f"(({lo}) + {v})")
            else:
                lidx.append(self.ex(x))
        sect = [v for v, _ in loops]
        self.emit("{")
        self.ind += 1
        for v, cnt in reversed(loops):       # first Fortran dimension innermost
            self.emit(f"for (int {v} = 0; {v} < {cnt}; {v}++)")
        self.emit(f"  {arr.elem(lidx)} = {self.ex(rhs, sect)};")
        self.ind -= 1
        self.emit("}")

    def call(self, f, a):
        sig = self.gen.sigs.get(f)
        out = []
        for k, x in enumerate(a):
            x = x.strip()
            want_array = sig is not None and k < len(sig) and sig[k]
            mm = re.match(r"^(\w+)\s*\((.*)\)$", x)
            if want_array and mm and mm.group(1) in self.scope:
                arr = self.scope[mm.group(1)]
                idx = [self.ex(y) for y in _split_top(mm.group(2))]
                # (an element of a rank-r array handed on with fewer subscripts cannot occur in Fortran)
                out.append(f"&{arr.elem(idx)}")
            elif want_array and x in self.scope:
                out.append(self.scope[x].cname)
            else:
                out.append(self.ex(x))
        self.emit(f"f_{f}({', '.join(out)});")

    def close_do(self):
        self.do_stack.pop()
        self.ind -= 1
        self.emit("}")

    def stmt(self, st):
        if st.startswith("$omp"):
            m = re.match(r"^\$omp\s+parallel\s+do\b(.*)$", st)
            if m and self.gen.openmp:
                cl = m.group(1)
                pv = re.search(r"private\s*\(([^)]*)\)", cl)
                sc = re.search(r"schedule\s*\(\s*static\s*,\s*(\w+)\s*\)", cl)
                s = "#pragma omp parallel for"
                if pv:
                    s += " private(" + ", ".join("f_" + x.strip() for x in pv.group(1).split(",")) + ")"
                if sc:
                    s += f" schedule(static, f_{sc.group(1)})"
                self.lines.append(s)
            return
        m = re.match(r"^if\s*\((.*)\)\s*go\s*to\s*(\d+)$", st)
        if m:
            self.emit(f"if ({self.ex(m.group(1))}) goto L{m.group(2)};")
            return
        m = re.match(r"^go\s*to\s*(\d+)$", st)
        if m:
            self.emit(f"goto L{m.group(1)};")
            return
        if re.match(r"^(where|forall|select|cycle|exit)\b", st):
            raise NotImplementedError(st)
        if _IO.match(st) or st == "continue":
            self.emit(";")
            return
        if _parse_decl(st) or re.match(r"^parameter\s*\(", st) or fx._DECL.match(st):
            return
        if st == "return":
            self.emit("return;")
            return
        if st.startswith("stop"):
            self.emit("f_stop_(); return;")
            return
        m = re.match(r"^do\s+(?:(\d+)\s+)?(\w+)\s*=\s*(.*)$", st)
        if m:
            parts = _split_top(m.group(3))
            v = f"f_{m.group(2)}"
            self.nloop += 1
            if len(parts) == 2:
                self.emit(f"for ({v} = ({self.ex(parts[0])}); {v} <= ({self.ex(parts[1])}); {v}++) {{")
            else:
                b, c = f"_b{self.nloop}", f"_c{self.nloop}"
                self.emit(f"const int {b} = ({self.ex(parts[1])}), {c} = ({self.ex(parts[2])});")
                self.emit(f"for ({v} = ({self.ex(parts[0])}); {c} > 0 ? {v} <= {b} : {v} >= {b}; {v} += {c}) {{")
            self.ind += 1
            self.do_stack.append(int(m.group(1)) if m.group(1) else None)
            return
        if re.match(r"^end\s*do$", st):
            self.close_do()
            return
        m = re.match(r"^(else\s*if|elseif|if)\s*\(", st)
        if m and m.group(1) == "if" and "if" in self.scope:
            e0 = _match_paren(st, st.index("("))
            if re.match(r"^\s*=[^=]", st[e0 + 1:]):
                m = None          # an array called `if` (bigrid.F90: indxi) is being assigned
        if m:
            start = st.index("(", m.end() - 1)
            end = _match_paren(st, start)
            cond, rest = st[start + 1:end], st[end + 1:].strip()
            if rest == "then":
                if m.group(1) != "if":
                    self.ind -= 1
                    self.emit(f"}} else if ({self.ex(cond)}) {{")
                else:
                    self.emit(f"if ({self.ex(cond)}) {{")
                self.ind += 1
            else:
                self.emit(f"if ({self.ex(cond)}) {{")
                self.ind += 1
                self.stmt(rest)
                self.ind -= 1
                self.emit("}")
            return
        if st == "else":
            self.ind -= 1
            self.emit("} else {")
            self.ind += 1
            return
        if re.match(r"^end\s*if$", st):
            self.ind -= 1
            self.emit("}")
            return
        m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", st)
        if m:
            f, a = m.group(1), (_split_top(m.group(2)) if m.group(2) else [])
            if f in self.gen.skip:
                self.emit(";")
                return
            self.call(f, a)
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
        m = re.match(r"^(\w+)\s*\((.*)\)$", lhs)
        if m:
            n, idx = m.group(1), _split_top(m.group(2))
            if n not in self.scope:
                if all(re.match(r"^[a-z_]\w*$", i) for i in idx):       # a statement function
                    self.funcs.add(n)
                    rt = self.decl.get(n, ("double",))[0]
                    ps = ", ".join(f"{self.decl.get(i, ('double',))[0]} f_{i}" for i in idx)
                    self.emit(f"{rt} f_{n}({ps}) {{ return {self.ex(rhs)}; }}")
                    return
                raise NotImplementedError(f"assignment to unknown array {n}: {st}")
            if any(":" in i for i in idx):
                self.array_assign(n, idx, rhs)
            else:
                self.emit(f"{self.scope[n].elem([self.ex(i) for i in idx])} = {self.ex(rhs)};")
        elif lhs in self.scope:
            self.array_assign(lhs, None, rhs)
        else:
            self.emit(f"f_{lhs} = {self.ex(rhs)};")

    def source(self):
        self.saved = set()
        for _, st in self.body:
            m = re.match(r"^save\s+(.*)$", st)
            if m:
                self.saved |= {x.strip() for x in m.group(1).split(",")}
        self.lines = []
        self.declarations()
        decl_lines, self.lines = self.lines, []
        # statement functions are part of the declarations: they come out in text order among the statements below
        inner = []
        for iname, iargs, ibody in self.internals:
            u = CUnit(self.gen, iname, iargs, ibody, host=self)
            self.gen.sigs[iname] = [u.is_array_dummy(a) for a in iargs]
        skipping = 0
        for lab, st in self.body:
            if skipping:
                if re.match(r"^if\s*\(.*\)\s*then$", st):
                    skipping += 1
                elif re.match(r"^end\s*if$", st):
                    skipping -= 1
                continue
            if any(r.search(st) for r in self.gen.drop) and re.match(r"^if\s*\(.*\)\s*then$", st):
                skipping = 1
                continue
            if lab is not None and not (self.do_stack and self.do_stack[-1] == lab):
                self.lines.append(f"L{lab}: ;")
            self.stmt(st)
            while lab is not None and self.do_stack and self.do_stack[-1] == lab:
                self.close_do()
        # the statement functions (one-line nested functions) must precede the internal procedures that use them
        stf = [l for l in self.lines if re.match(r"^\s*(double|int) f_\w+\(.*\) \{ return .*; \}$", l)]
        rest = [l for l in self.lines if l not in stf]
        for iname, iargs, ibody in self.internals:
            u = CUnit(self.gen, iname, iargs, ibody, host=self)
            u.funcs |= self.funcs
            u.ind = 2
            inner += ["  auto " + u.header() + ";"]
        for iname, iargs, ibody in self.internals:
            u = CUnit(self.gen, iname, iargs, ibody, host=self)
            u.funcs |= self.funcs
            u.scope.update(self.scope)
            u.ind = 2
            body = u.source()
            inner += ["  " + u.header() + " {"] + body + ["  }"]
        if self.host is not None:
            return decl_lines + stf + rest
        return [self.header() + " {"] + decl_lines + stf + inner + rest + ["}"]


class Generator:
    """C source for a set of units; `env`: the interpreter's environment (module variables)"""

    def __init__(self, env, skip=(), drop=(), openmp=False):
        self.env, self.skip, self.openmp = env, set(skip), openmp
        self.drop = [re.compile(r) for r in drop]
        self.modarr, self.modsc, self.sigs, self.units = {}, {}, {}, []
        for k, v in env.items():
            if isinstance(v, FArray):
                ct = "int32_t" if v.isint else "double"
                self.modarr[k] = Arr(f"f_{k}", ct, [f"f_{k}_lo[{d}]" for d in range(v.rank)],
                                     [f"f_{k}_n[{d}]" for d in range(v.rank)])
            elif isinstance(v, bool) or isinstance(v, (int, np.integer)):
                self.modsc[k] = "int"
            elif isinstance(v, float):
                self.modsc[k] = "double"

    def add(self, path, name, defines=("RELO",)):
        stmts = fx.load_source(path, defines, keep_omp=True)
        args, body, internals = fx.extract_unit(stmts, name)
        u = CUnit(self, name.lower(), args, body, internals)
        self.sigs[name.lower()] = [u.is_array_dummy(a) for a in args]
        self.units.append(u)
        return u

    def source(self):
        out = [PRELUDE]
        for k, ct in sorted(self.modsc.items()):
            out.append(f"{ct} f_{k};")
        for k, a in sorted(self.modarr.items()):
            out.append(f"{a.ctype} *f_{k}; int f_{k}_lo[{a.rank}], f_{k}_n[{a.rank}];")
        protos, bodies = [], []
        for u in self.units:
            protos.append(u.header() + ";")
        for u in self.units:
            bodies += u.source() + [""]
        return "\n".join(out + protos + [""] + bodies) + "\n"


class _Manifest:
    """what bind() needs of a Generator, read back from <library>.json where the reference tree is absent"""

    def __init__(self, d):
        self.modsc = d["modsc"]
        self.modarr = {k: Arr(f"f_{k}", ct, [None] * rank, [None] * rank) for k, (ct, rank) in d["modarr"].items()}


class Library:
    @classmethod
    def prebuilt(cls, so_path):
        """a library built earlier (in the container that has /root/reference), with its manifest"""
        import json
        self = cls.__new__(cls)
        self.gen = _Manifest(json.load(open(so_path[:-3] + ".json")))
        self.lib = C.CDLL(so_path)
        self.keep = []
        return self

    def __init__(self, gen, so_path, flags=("-O2", "-ffp-contract=off")):
        import json
        self.gen = gen
        os.makedirs(os.path.dirname(so_path), exist_ok=True)
        c_path = so_path[:-3] + ".c"
        src = "/* " + " ".join(flags) + " */\n" + gen.source()
        json.dump(dict(modsc=gen.modsc, modarr={k: (a.ctype, a.rank) for k, a in gen.modarr.items()}),
                  open(so_path[:-3] + ".json", "w"))
        if not (os.path.exists(c_path) and open(c_path).read() == src and os.path.exists(so_path)):
            open(c_path, "w").write(src)
            gcc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"   # (the image's CC wrapper lacks libgomp.spec)
            cmd = [gcc, "-std=gnu11", "-shared", "-fPIC", "-w", *flags, *(["-fopenmp"] if gen.openmp else []),
                   "-o", so_path, c_path, "-lm"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                raise RuntimeError("gcc failed:\n" + r.stderr[:4000])
        self.lib = C.CDLL(so_path)
        self.keep = []

    def bind(self, env):
        """module variables of the interpreter's environment -> the globals of the library (arrays are shared)"""
        for k, ct in self.gen.modsc.items():
            if k in env:
                (C.c_int if ct == "int" else C.c_double).in_dll(self.lib, f"f_{k}").value = env[k]
        for k, a in self.gen.modarr.items():
            if k not in env:
                continue
            v = env[k]
            assert v.a.flags["C_CONTIGUOUS"] and v.a.dtype == (np.int32 if a.ctype == "int32_t" else np.float64), (k, v.a.dtype)
            C.c_void_p.in_dll(self.lib, f"f_{k}").value = v.a.ctypes.data
            lo = (C.c_int * a.rank).in_dll(self.lib, f"f_{k}_lo")
            nn = (C.c_int * a.rank).in_dll(self.lib, f"f_{k}_n")
            for d in range(a.rank):
                lo[d] = v.lo[d]
                nn[d] = v.a.shape[a.rank - 1 - d]
            self.keep.append(v.a)

    def pull(self, env):
        """module scalars the text assigned (nreg in bigrid ...) back into the environment"""
        for k, ct in self.gen.modsc.items():
            if k in env:
                val = (C.c_int if ct == "int" else C.c_double).in_dll(self.lib, f"f_{k}").value
                env[k] = bool(val) if isinstance(env[k], bool) else val

    def call(self, name, *args):
        f = getattr(self.lib, f"f_{name}")
        conv = []
        for a in args:
            if isinstance(a, FArray):
                conv.append(C.c_void_p(a.a.ctypes.data))
            elif isinstance(a, np.ndarray):
                conv.append(C.c_void_p(a.ctypes.data))
            elif isinstance(a, float):
                conv.append(C.c_double(a))
            else:
                conv.append(C.c_int(int(a)))
        f.restype = None
        f(*conv)
