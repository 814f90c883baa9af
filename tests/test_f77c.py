"""Unit tests of oracle/f77c.py, the fixed-form F77 -> C translator that lets the reference's own statements run here
(oracle/ref_build.py).  Small Fortran programs written for this test (not taken from the reference) exercise the language
features the reference's hot path relies on; each is translated, compiled with gcc -O2 -ffp-contract=off and compared with a
known answer computed independently in Python.  The strongest check of the translator remains tests/test_ref_pins.py: an
independently hand-written restatement agrees with the translated reference bit for bit."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

SRC = r"""
      subroutine dots(a,b,n,s)
c     implicit typing with REAL = 8 bytes, DO loop with a labelled CONTINUE, left-to-right accumulation
      real a(1),b(1)
      s = 0.
      do 10 i=1,n
         s = s + a(i)*b(i)
 10   continue
      return
      end
c-----------------------------------------------------------------------
      subroutine fill2d(a,m,n)
c     adjustable column-major array, nested loops, integer arithmetic, mixed-mode expression
      integer m,n
      real a(m,n)
      do j=1,n
      do i=1,m
         a(i,j) = 10*i + j + (i/2)*0.5 + mod(i*j,3)
      enddo
      enddo
      return
      end
c-----------------------------------------------------------------------
      subroutine com_put(x,k)
c     COMMON storage shared between routines with different views + EQUIVALENCE
      common /blk/ v(4),iv(2)
      real w(2)
      equivalence (v(3),w(1))
      v(1) = x
      v(2) = 2*x
      w(1) = -x
      w(2) = x**2
      iv(1) = k
      iv(2) = k*k
      return
      end
      subroutine com_get(out,iout)
      common /blk/ a(2),b(2),ia,ib
      real out(4)
      integer iout(2)
      out(1) = a(1)
      out(2) = a(2)
      out(3) = b(1)
      out(4) = b(2)
      iout(1) = ia
      iout(2) = ib
      return
      end
c-----------------------------------------------------------------------
      integer function counter()
c     SAVE + DATA: state survives between calls
      integer icalld
      save    icalld
      data    icalld /0/
      icalld = icalld + 1
      counter = icalld
      return
      end
c-----------------------------------------------------------------------
      subroutine branchy(x,n,out)
c     block IF / ELSEIF, arithmetic intrinsics, GOTO out of a loop, logical operators, PARAMETER constants
      parameter (lim=5, half=0.5)
      real x(n)
      logical big
      out = 0.
      do i=1,n
         big = x(i).gt.half .and. .not.(x(i).ge.2.)
         if (big) then
            out = out + sqrt(abs(x(i)))
         elseif (x(i).lt.0) then
            out = out - min(1.,max(-1.,x(i)))
         else
            out = out + sign(half,x(i)-1.)
         endif
         if (i.ge.lim) goto 20
      enddo
 20   continue
      out = out + i
      return
      end
c-----------------------------------------------------------------------
      subroutine named(name,val)
c     CHARACTER dummy with hidden length, comparison against a blank-padded literal
      character*4 name
      val = 0.
      if (name.eq.'PRES') val = 1.
      if (name.eq.'VELX') val = 2.
      if (name.eq.'bp5 ') val = 3.
      return
      end
c-----------------------------------------------------------------------
      subroutine caller(a,n,s)
c     by-reference calls: array element as the start of a sub-array, expression temporaries, function result
      real a(n)
      integer counter
      m = n/2
      call dots(a(m+1),a(m+1),n-m,s)
      k = counter()
      s = s + k
      return
      end
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    import f77c
    d = tmp_path_factory.mktemp("f77c")
    src = d / "unit.f"
    src.write_text(SRC)
    tr = f77c.Translator(str(d), [str(d)], {}, defines=[])
    tr.add_file(str(src))
    code, missing = tr.translate(["dots", "fill2d", "com_put", "com_get", "counter", "branchy", "named", "caller"])
    assert not missing and not tr.failed
    cfile, so = d / "unit.c", d / "libunit.so"
    cfile.write_text(code)
    r = subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-w", "-fPIC", "-shared", "-o", str(so), str(cfile), "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return C.CDLL(str(so)), tr


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_do_loop_and_accumulation_order(lib):
    L, _ = lib
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(1000), rng.standard_normal(1000)
    s = C.c_double(0)
    L.dots_(_p(a), _p(b), C.byref(C.c_int(1000)), C.byref(s))
    ref = 0.0
    for x, y in zip(a, b):              # the same left-to-right order, no FMA contraction
        ref = ref + x * y
    assert s.value == ref


def test_column_major_arrays_and_integer_arithmetic(lib):
    L, _ = lib
    m, n = 5, 4
    a = np.zeros((m, n), order="F")
    L.fill2d_(_p(a), C.byref(C.c_int(m)), C.byref(C.c_int(n)))
    for i in range(1, m + 1):
        for j in range(1, n + 1):
            assert a[i - 1, j - 1] == 10 * i + j + (i // 2) * 0.5 + (i * j) % 3


def test_common_blocks_are_raw_storage_with_per_routine_views(lib):
    L, tr = lib
    L.com_put_(C.byref(C.c_double(1.5)), C.byref(C.c_int(7)))
    out, iout = np.zeros(4), np.zeros(2, dtype=np.int32)
    L.com_get_(_p(out), _p(iout))
    assert out.tolist() == [1.5, 3.0, -1.5, 2.25] and iout.tolist() == [7, 49]
    assert tr.common_size["blk"] == 4 * 8 + 2 * 4


def test_save_and_data(lib):
    L, _ = lib
    L.counter_.restype = C.c_int
    first = L.counter_()
    assert [L.counter_() for _ in range(3)] == [first + 1, first + 2, first + 3]


def test_control_flow_and_intrinsics(lib):
    L, _ = lib
    x = np.array([0.7, -3.0, 5.0, 1.5, 0.2, 9.0, 9.0])
    out = C.c_double(0)
    L.branchy_(_p(x), C.byref(C.c_int(len(x))), C.byref(out))
    ref = np.sqrt(0.7) + 1.0 + 0.5 + np.sqrt(1.5) - 0.5       # five entries, then the GOTO leaves with i = 5
    assert abs(out.value - (ref + 5)) <= 1e-15
    out2 = C.c_double(0)
    L.branchy_(_p(x), C.byref(C.c_int(3)), C.byref(out2))     # loop runs to completion: i = n + 1 afterwards
    assert abs(out2.value - (np.sqrt(0.7) + 1.0 + 0.5 + 4)) <= 1e-15


def test_character_arguments_with_hidden_length(lib):
    L, _ = lib
    for name, want in ((b"PRES", 1.0), (b"VELX", 2.0), (b"bp5", 3.0), (b"TEMP", 0.0)):
        v = C.c_double(-1)
        L.named_(C.c_char_p(name.ljust(4)), C.byref(v), C.c_long(4))
        assert v.value == want, name


def test_calls_pass_addresses(lib):
    L, _ = lib
    L.counter_.restype = C.c_int
    a = np.arange(1.0, 9.0)
    k0 = L.counter_()
    s = C.c_double(0)
    L.caller_(_p(a), C.byref(C.c_int(8)), C.byref(s))
    assert s.value == float(np.sum(a[4:] ** 2) + k0 + 1)
