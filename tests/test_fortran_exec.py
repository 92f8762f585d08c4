"""oracle/fortran_exec.py on its own: the pin of the oracle (tests/test_reference_text.py) rests on this translator
executing Fortran the way a compiler does, so every language rule the reference files rely on is checked here on
small programs with known answers - integer arithmetic, evaluation order, loops, branches, array bounds and
sections, sequence association, statement functions, internal procedures, cpp."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fortran_exec as fx  # noqa: E402


def run(tmp_path, text, name, env=None, defines=("RELO",), **kw):
    f = tmp_path / "unit.F90"
    f.write_text(text)
    env = {} if env is None else env
    fx.compile_unit(str(f), name, env, defines=defines, **kw)
    return env


def test_integer_arithmetic_follows_fortran(tmp_path):
    env = run(tmp_path, """
      subroutine t(out)
      integer out(12)
      integer i, j
      i = 7 ; j = -7
      out(1) = i/2
      out(2) = j/2
      out(3) = mod(i,3)
      out(4) = mod(j,3)
      out(5) = mod(i,-3)
      out(6) = 2**5
      out(7) = (-2)**3
      out(8) = nint(2.5)
      out(9) = nint(-2.5)
      out(10)= int(-2.7)
      out(11)= i/2*2
      out(12)= 2*i/2
      end subroutine t
    """, "t")
    out = fx.FArray.zeros(((1, 12),), dtype=np.int64)
    env["t"](out)
    assert out.a.tolist() == [3, -3, 1, -1, 1, 32, -8, 3, -3, -2, 6, 7]


def test_real_arithmetic_order_and_intrinsics(tmp_path):
    env = run(tmp_path, """
      subroutine t(out, a, b, c)
      real out(10), a, b, c
      out(1) = a+b+c
      out(2) = a+(b+c)
      out(3) = a*b*c
      out(4) = a*(b*c)
      out(5) = -a**2
      out(6) = sign(3.0,-0.0) + sign(2.0, 0.0)
      out(7) = max(0., -0.)
      out(8) = max(-0., 0.)
      out(9) = 1./3.
      out(10)= 2.e0**0.5 + 1.d0
      end subroutine t
    """, "t")
    out = fx.FArray.zeros(((1, 10),))
    a, b, c = 1.0e16, -1.0e16, 1.0
    env["t"](out, a, b, c)
    o = out.a
    assert o[0] == (a + b) + c == 1.0 and o[1] == a + (b + c) == 0.0
    x, y, z = 1.1, 1.3, 1.7
    env["t"](out, x, y, z)
    assert o[2] == (x * y) * z and o[3] == x * (y * z)
    assert o[4] == -(x * x)
    assert o[5] == -3.0 + 2.0
    assert o[6] == 0.0 and not math.copysign(1.0, o[6]) < 0          # gfortran: the first argument stays on a tie
    assert o[7] == 0.0 and math.copysign(1.0, o[7]) < 0
    assert o[8] == 1.0 / 3.0 and o[9] == math.sqrt(2.0) + 1.0


def test_loops_branches_labels_and_goto(tmp_path):
    env = run(tmp_path, """
      subroutine t(out, n, flag)
      integer out(8), n, flag
      integer i, j, k, s
      s = 0
      do i= n,1,-2
        s = s + i
      enddo
      out(1) = s
      out(2) = i
      s = 0
      do i= 5,4
        s = s + 1
      enddo
      out(3) = s
      s = 0
      do 10 j=1,3
      do 10 k=1,2
        s = s + j*k
 10   continue
      out(4) = s
      if     (n.gt.8) then
        out(5) = 1
      elseif (n.gt.6 .and. .not.(n.eq.8)) then
        out(5) = 2
      else
        out(5) = 3
      endif
      if (n >= 7) out(6) = 6
      out(7) = 0
      if (flag.eq.0) go to 20
      out(7) = 7
      do i=1,2
        out(7) = out(7) + i
      enddo
 20   continue
      out(8) = 8
      end subroutine t
    """, "t")
    out = fx.FArray.zeros(((1, 8),), dtype=np.int64)
    env["t"](out, 7, 0)
    assert out.a.tolist() == [7 + 5 + 3 + 1, -1, 0, 18, 2, 6, 0, 8]
    env["t"](out, 7, 1)
    assert out.a[6] == 10


def test_array_bounds_sections_and_sequence_association(tmp_path):
    env = dict(nbdy=2, idm=3, jdm=2)
    run(tmp_path, """
      subroutine fill(a, l1, ld, v)
      integer l1, ld
      real a(1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy,ld), v
      integer i, j, k
      do k= l1,ld
        do j= 1,jdm
          do i= 1,idm
            a(i,j,k) = v + 100*k + 10*j + i
          enddo
        enddo
      enddo
      end subroutine fill
      subroutine t(f, g2)
      real f(1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy,2,2)
      real g2(1-nbdy:idm+nbdy,1-nbdy:jdm+nbdy)
      f(:,:,:,:) = -1.0
      call fill(f(1-nbdy,1-nbdy,2,1), 1, 2, 0.5)
      g2 = 3.0
      call fill(g2(1-nbdy,1-nbdy), 1, 1, 0.25)
      g2(1-nbdy,:) = 9.0
      f(0,0,1,1) = f(1,1,2,1) + f(3,2,1,2)
      end subroutine t
    """, "fill", env)
    run(tmp_path, open(tmp_path / "unit.F90").read(), "t", env, callee_ranks={"fill": (3, None, None, None)})
    f = fx.FArray.zeros(((-1, 5), (-1, 4), (1, 2), (1, 2)))
    g2 = fx.FArray.zeros(((-1, 5), (-1, 4)))
    env["t"](f, g2)
    # the callee saw slabs (2,1) and (1,2) of f as its k = 1, 2: storage is [t][k][j][i]
    assert f.a[0, 1, 2, 2] == 0.5 + 100 + 10 + 1          # f(1,1,2,1): callee k=1
    assert f.a[1, 0, 3, 4] == 0.5 + 200 + 20 + 3          # f(3,2,1,2): callee k=2
    assert f.a[1, 1, 2, 2] == -1.0                        # f(1,1,2,2): beyond the callee's ld
    assert f.a[0, 0, 2, 2] == -1.0                        # f(1,1,1,1): before the element passed
    assert f.a[0, 0, 1, 1] == f.a[0, 1, 2, 2] + f.a[1, 0, 3, 4]
    assert g2.a[2, 2] == 0.25 + 100 + 10 + 1 and g2.a[0, 1] == 3.0
    assert (g2.a[:, 0] == 9.0).all()


def test_statement_functions_parameters_and_cpp(tmp_path):
    (tmp_path / "fns.h").write_text("""
      real sq, harm, x, y
      real, parameter :: half=0.5
      real, parameter, dimension(3) :: cc = (/ 1.0, 2.0, &
                                               4.0 /)
      sq(x)=x*x
      harm(x,y)=2.0*x*y/(x+y)
""")
    text = """
#if defined(BIG)
#define SCALE 10.0
#elif defined(SMALL) || defined(TINY)
#define SCALE 0.1
#else
#define SCALE 1.0
#endif
#define SEA_P ip(i).ne.0
      subroutine t(out, ip)
      real out(4)
      integer ip(4)
      integer i
      integer, parameter :: n=4
# include "fns.h"
      do i= 1,n
        if (SEA_P) then
          out(i) = SCALE*(sq(real(i)) + harm(cc(1),cc(3))*half)   ! a comment with 'quotes' and ; in it
        else
          out(i) = -1.0 &
                 & -1.0
        endif
      enddo
      end subroutine t
"""
    ip = fx.FArray(np.array([1, 0, 1, 1], dtype=np.int64), (1,))
    for defs, scale in ((("BIG",), 10.0), (("TINY",), 0.1), ((), 1.0)):
        env = run(tmp_path, text, "t", defines=defs)
        out = fx.FArray.zeros(((1, 4),))
        env["t"](out, ip)
        want = [scale * (i * i + (2.0 * 1.0 * 4.0 / (1.0 + 4.0)) * 0.5) if m else -2.0 for i, m in zip((1.0, 2.0, 3.0, 4.0), ip.a)]
        assert out.a.tolist() == want


def test_internal_procedures_share_the_host_but_keep_their_locals(tmp_path):
    env = dict(total=0.0)
    run(tmp_path, """
      subroutine t(out, n)
      real out(3)
      integer n
      integer i, k
      real acc
      acc = 0.0
      do k= 1,n
        call add(real(k))
      enddo
      out(1) = acc
      out(2) = k
      out(3) = i
      total = acc
      return
      contains
      subroutine add(v)
      real v
      integer k
      do k= 1,2
        acc = acc + v
      enddo
      i = 42
      end subroutine add
      end subroutine t
    """, "t", env)
    out = fx.FArray.zeros(((1, 3),))
    env["t"](out, 3)
    assert out.a.tolist() == [2.0 * (1 + 2 + 3), 4.0, 42.0]
    assert env["total"] == 12.0                              # a module variable assigned by the unit


def test_unsupported_constructs_raise(tmp_path):
    for body in ("where (a.gt.0.) a = 0.", "go to 10", "select case (i)"):
        with pytest.raises(NotImplementedError):
            run(tmp_path, f"""
      subroutine t(a)
      real a(3)
      {body}
      end subroutine t
    """, "t")


def test_the_two_backends_agree_on_random_expressions(tmp_path):
    """oracle/fortran_exec.py (Python) and oracle/fortran_to_c.py (C, gcc) share the tokenizer and nothing else:
    300 random Fortran expressions - mixed integer / real arithmetic, parentheses, unary minus, integer powers,
    max / min / abs / sign / mod / int / real / sqrt, relational and logical operators in an if - must come out
    bit-identical from both"""
    import random
    import shutil
    if not (shutil.which("gcc") or os.access("/usr/bin/gcc", os.X_OK)):
        pytest.skip("no gcc")
    import fortran_to_c as f2c
    rnd = random.Random(20261018)
    reals, ints = ["a", "b", "c"], ["i", "j"]

    def rexpr(d):
        if d <= 0 or rnd.random() < 0.2:
            return rnd.choice(reals + ["1.5", "0.25", "3.e-2", "2.d0", "7."])
        k = rnd.randrange(12)
        x, y = rexpr(d - 1), rexpr(d - 1)
        if k < 4:
            return f"({x} {rnd.choice('+-*/')} {y})"
        if k == 4:
            return f"{x}{rnd.choice('+-*')}{y}"
        if k == 5:
            return f"(-{x})"
        if k == 6:
            return f"{rnd.choice(['max', 'min'])}({x},{y},{rexpr(d - 1)})"
        if k == 7:
            return rnd.choice([f"abs({x})", f"sqrt(abs({x}))"])
        if k == 8:
            return f"sign({x},{y})"
        if k == 9:
            return f"({x})**{rnd.choice(['2', '3', '(-2)'])}"
        if k == 10:
            return f"real({iexpr(d - 1)})"
        return f"({x} * {iexpr(d - 1)})"

    def iexpr(d):
        if d <= 0 or rnd.random() < 0.3:
            return rnd.choice(ints + ["2", "3", "7"])
        k = rnd.randrange(6)
        x, y = iexpr(d - 1), iexpr(d - 1)
        if k < 3:
            return f"({x} {rnd.choice('+-*')} {y})"
        if k == 3:
            return f"({x}/(abs({y})+1))"
        if k == 4:
            return f"mod({x},abs({y})+2)"
        return f"max({x},{y})"
    exprs = []
    while len(exprs) < 300:
        exprs.append(rexpr(4))
    body = "\n".join(f"      if (({rexpr(2)}) .lt. ({rexpr(2)}) .or. .not. (({iexpr(2)}) .ge. ({iexpr(2)}))) then\n"
                     f"        out({k + 1}) = {e}\n      else\n        out({k + 1}) = -({e})\n      endif" for k, e in enumerate(exprs))
    text = f"""
      subroutine t(out, a, b, c, i, j)
      real out({len(exprs)}), a, b, c
      integer i, j
{body}
      end subroutine t
"""
    f = tmp_path / "unit.F90"
    f.write_text(text)
    env = {}
    fx.compile_unit(str(f), "t", env)
    gen = f2c.Generator({})
    gen.add(str(f), "t")
    lib = f2c.Library(gen, str(tmp_path / "libt.so"))
    for a, b, c, i, j in ((1.25, -2.5, 3.0e-3, 3, -4), (-7.0, 0.1, 1.0e10, -2, 5), (0.0, -0.0, 1.0, 0, 1)):
        po = fx.FArray.zeros(((1, len(exprs)),))
        co = fx.FArray.zeros(((1, len(exprs)),))
        with np.errstate(all="ignore"):
            env["t"](po, a, b, c, i, j)
        lib.call("t", co, a, b, c, i, j)
        same = (po.a.view(np.uint64) == co.a.view(np.uint64)) | (np.isnan(po.a) & np.isnan(co.a))
        bad = np.flatnonzero(~same)
        assert bad.size == 0, [(exprs[k], po.a[k], co.a[k]) for k in bad[:5]]
