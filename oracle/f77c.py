"""f77c -- a small fixed-form FORTRAN-77 -> C translator, TEST INFRASTRUCTURE ONLY.

Purpose: neither a Fortran compiler nor gslib exists in the authoring image, so the reference (Nek5000, F77) cannot be
built with its own toolchain.  This translator reads the reference's *.f / include files WHERE THEY LIE under
/root/reference (never copied into the repo), emits C for the routines the hot path uses into oracle/_ref/ (git-ignored)
and that C is compiled into oracle/_ref/libnekref_*.so by oracle/ref_build.py.  The result is the reference's own
statements executing here: it pins the hand-written restatement in oracle/nek_oracle.c / oracle/hsmg.py, and it is the
generator of the golden vectors under tests/golden/.  Nothing under nek5000_b200/ may import it.

Semantics honoured (what the needed routines use): fixed-form source with cpp conditionals, INCLUDE, implicit typing
with the reference's `-fdefault-real-8 -fdefault-double-8` (REAL = 8 bytes, INTEGER/LOGICAL = 4), PARAMETER constant
folding, COMMON blocks as raw storage (every routine lays its own view over the block, as Fortran does), EQUIVALENCE,
SAVE/DATA, adjustable and assumed-size dummy arrays with arbitrary lower bounds, by-reference argument passing (array
elements pass their address, expressions pass a temporary), hidden CHARACTER lengths appended gfortran-style, DO loops
with labelled/shared terminal statements, block IF, GOTO, statement order of every floating-point expression
(no re-association: the C is compiled with -O2 -ffp-contract=off, the reference default being -O2 on baseline x86-64).
I/O statements (write/print/read/format/open/close) are dropped.  Anything unsupported aborts translation loudly.
"""
from __future__ import annotations

import json
import os
import re
import sys

# ----------------------------------------------------------------------------------------------------------------------
# source reading: cpp conditionals, fixed-form continuation, includes
# ----------------------------------------------------------------------------------------------------------------------


class Stmt:
    __slots__ = ("label", "text", "file", "line", "toks")

    def __init__(self, label, text, file, line):
        self.label, self.text, self.file, self.line, self.toks = label, text, file, line, None

    def __repr__(self):
        return f"<{os.path.basename(self.file)}:{self.line} {self.label or ''} {self.text}>"


def strip_comment(s: str) -> str:
    """Removes a trailing `! comment` (outside character literals)."""
    q = None
    for i, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return s[:i]
    return s


class Reader:
    def __init__(self, include_dirs, include_map=None, defines=()):
        self.include_dirs = list(include_dirs)
        self.include_map = dict(include_map or {})
        self.defines = set(defines)
        self._cache = {}

    def find_include(self, name):
        name = self.include_map.get(name, name)
        if os.path.isabs(name) and os.path.exists(name):
            return name
        for d in self.include_dirs:
            p = os.path.join(d, name)
            if os.path.exists(p):
                return p
        raise FileNotFoundError(f"include file {name!r} not found in {self.include_dirs}")

    def read(self, path):
        """Returns the list of Stmt of a file with its includes expanded inline."""
        if path in self._cache:
            return self._cache[path]
        raw = open(path, errors="replace").read().split("\n")
        # cpp conditionals
        lines = []
        stack = []  # (active_before, taken)
        active = True
        for ln, s in enumerate(raw, 1):
            if s.startswith("#"):
                d = s[1:].strip().split()
                if not d:
                    continue
                key = d[0]
                if key in ("ifdef", "ifndef"):
                    cond = (d[1] in self.defines) == (key == "ifdef")
                    stack.append((active, cond))
                    active = active and cond
                elif key == "if":
                    m = re.match(r"defined\s*\(?\s*(\w+)\s*\)?$", " ".join(d[1:]))
                    cond = bool(m and m.group(1) in self.defines)
                    stack.append((active, cond))
                    active = active and cond
                elif key == "else":
                    prev, cond = stack[-1]
                    active = prev and not cond
                elif key == "endif":
                    prev, _ = stack.pop()
                    active = prev
                elif key == "define" and active:
                    self.defines.add(d[1])
                elif key == "undef" and active:
                    self.defines.discard(d[1])
                continue
            if active:
                lines.append((ln, s))
        # fixed form -> logical statements
        out = []
        cur = None
        for ln, s in lines:
            if not s.strip():
                continue
            c0 = s[0]
            if c0 in "cC*!dD":
                continue
            if "\t" in s[:6]:
                s = s.replace("\t", "      ", 1)
            s = s[:72]
            if s.lstrip().startswith("!"):
                continue
            s = strip_comment(s)
            if not s.strip():
                continue
            lab = s[:5].strip()
            cont = len(s) > 5 and s[5] not in " 0"
            body = s[6:]
            if cont and cur is not None and not lab:
                cur.text += body
                continue
            if cur is not None:
                out.append(cur)
            cur = Stmt(lab if lab else None, body, path, ln)
        if cur is not None:
            out.append(cur)
        # includes
        res = []
        for st in out:
            m = re.match(r"\s*include\s*['\"]([^'\"]+)['\"]\s*$", st.text, re.I)
            if m:
                res.extend(self.read(self.find_include(m.group(1).strip())))
            else:
                st.text = st.text.strip()
                if st.text:
                    res.append(st)
        self._cache[path] = res
        return res


# ----------------------------------------------------------------------------------------------------------------------
# tokenizer
# ----------------------------------------------------------------------------------------------------------------------

DOTOPS = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".and.": "&&", ".or.": "||",
          ".not.": "!", ".eqv.": "eqv", ".neqv.": "neqv", ".true.": "true", ".false.": "false"}
_dot_re = re.compile(r"\.(eq|ne|lt|le|gt|ge|and|or|not|eqv|neqv|true|false)\.", re.I)
_name_re = re.compile(r"[A-Za-z_$][A-Za-z0-9_$]*")
_num_re = re.compile(r"(\d+\.?\d*|\.\d+)([eEdD][+-]?\d+)?(_\w+)?")


def tokenize(s: str):
    toks = []
    i, n = 0, len(s)
    while i < n:
        ch = s[i]
        if ch in " \t":
            i += 1
            continue
        if ch in "'\"":
            j = i + 1
            buf = []
            while True:
                if j >= n:
                    raise SyntaxError(f"unterminated string in {s!r}")
                if s[j] == ch:
                    if j + 1 < n and s[j + 1] == ch:
                        buf.append(ch)
                        j += 2
                        continue
                    break
                buf.append(s[j])
                j += 1
            toks.append(("str", "".join(buf)))
            i = j + 1
            continue
        if ch == ".":
            m = _dot_re.match(s, i)
            if m:
                toks.append(("op", DOTOPS[m.group(0).lower()]))
                i = m.end()
                continue
        if ch.isdigit() or (ch == "." and i + 1 < n and s[i + 1].isdigit()):
            # integer or real; do not swallow the '.' of a following dot-operator (1.eq.2)
            j = i
            while j < n and s[j].isdigit():
                j += 1
            isreal = False
            if j < n and s[j] == "." and not _dot_re.match(s, j):
                isreal = True
                j += 1
                while j < n and s[j].isdigit():
                    j += 1
            m = re.compile(r"[eEdD][+-]?\d+").match(s, j)
            if m:
                isreal = True
                j = m.end()
            txt = s[i:j]
            m = re.compile(r"_\w+").match(s, j)
            if m:
                j = m.end()
            toks.append(("real" if isreal else "int", txt))
            i = j
            continue
        m = _name_re.match(s, i)
        if m:
            toks.append(("name", m.group(0).lower()))
            i = m.end()
            continue
        two = s[i:i + 2]
        if two in ("**", "//", "==", "/=", "<=", ">="):
            toks.append(("op", {"/=": "!=", "**": "**", "//": "//", "==": "==", "<=": "<=", ">=": ">="}[two]))
            i += 2
            continue
        if ch in "+-*/(),=:<>":
            toks.append(("op", ch))
            i += 1
            continue
        raise SyntaxError(f"bad character {ch!r} in {s!r}")
    return toks


# ----------------------------------------------------------------------------------------------------------------------
# expression parser -> AST tuples
#   ('int', text) ('real', text) ('str', s) ('log', 0|1) ('name', n) ('app', n, [args]) ('bin', op, a, b) ('un', op, a)
#   ('range', lo|None, hi|None)
# ----------------------------------------------------------------------------------------------------------------------


class Parser:
    def __init__(self, toks, where=""):
        self.t, self.i, self.where = toks, 0, where

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def at(self, kind, val=None):
        tok = self.peek()
        return tok[0] == kind and (val is None or tok[1] == val)

    def expect(self, kind, val=None):
        tok = self.next()
        if tok[0] != kind or (val is not None and tok[1] != val):
            raise SyntaxError(f"{self.where}: expected {val or kind}, got {tok} in {self.t}")
        return tok

    def done(self):
        return self.i >= len(self.t)

    # precedence climbing
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        a = self.p_or()
        while self.at("op", "eqv") or self.at("op", "neqv"):
            op = self.next()[1]
            a = ("bin", op, a, self.p_or())
        return a

    def p_or(self):
        a = self.p_and()
        while self.at("op", "||"):
            self.next()
            a = ("bin", "||", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.at("op", "&&"):
            self.next()
            a = ("bin", "&&", a, self.p_not())
        return a

    def p_not(self):
        if self.at("op", "!"):
            self.next()
            return ("un", "!", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_cat()
        if self.peek()[0] == "op" and self.peek()[1] in ("==", "!=", "<", "<=", ">", ">="):
            op = self.next()[1]
            a = ("bin", op, a, self.p_cat())
        return a

    def p_cat(self):
        a = self.p_add()
        while self.at("op", "//"):
            self.next()
            a = ("bin", "//", a, self.p_add())
        return a

    def p_add(self):
        if self.at("op", "-") or self.at("op", "+"):
            op = self.next()[1]
            a = self.p_mul()
            if op == "-":
                a = ("un", "-", a)
        else:
            a = self.p_mul()
        while self.at("op", "+") or self.at("op", "-"):
            op = self.next()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.at("op", "*") or self.at("op", "/"):
            op = self.next()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_prim()
        if self.at("op", "**"):
            self.next()
            # right associative; exponent may carry a unary sign
            if self.at("op", "-") or self.at("op", "+"):
                op = self.next()[1]
                b = self.p_pow()
                if op == "-":
                    b = ("un", "-", b)
            else:
                b = self.p_pow()
            a = ("bin", "**", a, b)
        return a

    def p_prim(self):
        tok = self.next()
        k, v = tok
        if k == "int":
            return ("int", v)
        if k == "real":
            return ("real", v)
        if k == "str":
            return ("str", v)
        if k == "op" and v == "true":
            return ("log", 1)
        if k == "op" and v == "false":
            return ("log", 0)
        if k == "op" and v == "(":
            e = self.expr()
            self.expect("op", ")")
            return ("paren", e)
        if k == "op" and v in "+-":
            a = self.p_prim()
            return ("un", "-", a) if v == "-" else a
        if k == "name":
            if self.at("op", "("):
                self.next()
                args = self.arglist()
                node = ("app", v, args)
                if self.at("op", "("):          # substring of an array element: a(i)(1:3)
                    self.next()
                    sub = self.arglist()
                    node = ("substr", node, sub[0])
                return node
            return ("name", v)
        raise SyntaxError(f"{self.where}: unexpected token {tok} in {self.t}")

    def arglist(self):
        args = []
        if self.at("op", ")"):
            self.next()
            return args
        while True:
            args.append(self.arg())
            if self.at("op", ","):
                self.next()
                continue
            self.expect("op", ")")
            return args

    def arg(self):
        # expression or range lo:hi with either side optional
        if self.at("op", ":"):
            self.next()
            hi = None if (self.at("op", ")") or self.at("op", ",")) else self.expr()
            return ("range", None, hi)
        if self.at("op", "*") and self.peek(1)[0] == "op" and self.peek(1)[1] in (")", ","):
            self.next()
            return ("star",)
        e = self.expr()
        if self.at("op", ":"):
            self.next()
            if self.at("op", "*"):
                self.next()
                return ("range", e, ("star",))
            hi = None if (self.at("op", ")") or self.at("op", ",")) else self.expr()
            return ("range", e, hi)
        return e


# ----------------------------------------------------------------------------------------------------------------------
# types
# ----------------------------------------------------------------------------------------------------------------------

CTYPE = {"i4": "int", "i8": "long long", "r8": "double", "r4": "float", "l4": "int", "ch": "char", "i2": "short",
         "i1": "signed char", "l1": "signed char"}
TSIZE = {"i4": 4, "i8": 8, "r8": 8, "r4": 4, "l4": 4, "ch": 1, "i2": 2, "i1": 1, "l1": 1}
RANK = {"l4": 0, "i1": 1, "i2": 1, "i4": 1, "i8": 2, "r4": 3, "r8": 4}

C_RESERVED = set("""auto break case char const continue default do double else enum extern float for goto if inline int long
register restrict return short signed sizeof static struct switch typedef union unsigned void volatile while main linux unix
y0 y1 yn j0 j1 jn gamma index time exp log sin cos tan pow sqrt abs fabs floor ceil round trunc fmod div free malloc
erf erfc signal remove rename exit abort rand stdin stdout stderr errno clock read write open close link""".split())


def cname(n: str) -> str:
    n = n.replace("$", "_S_")
    return n + "_v" if n in C_RESERVED else n


class Sym:
    def __init__(self, name):
        self.name = name
        self.typ = None        # 'i4','r8',... ; None -> implicit
        self.clen = None       # character length (int) or '*'
        self.dims = None       # list of (lo_ast, hi_ast|None)
        self.arg = False
        self.common = None     # (block, byte offset)
        self.param = None      # constant value (python int/float/str/bool)
        self.save = False
        self.data = False
        self.external = False
        self.eqv = None        # (blob name, byte offset)
        self.explicit = False
        self.is_func_result = False

    @property
    def is_array(self):
        return self.dims is not None


INTRINSICS = set("""abs iabs dabs sqrt dsqrt exp dexp log alog dlog log10 alog10 dlog10 sin dsin cos dcos tan dtan asin dasin
acos dacos atan datan atan2 datan2 sinh dsinh cosh dcosh tanh dtanh max min amax1 amin1 max0 min0 dmax1 dmin1 amax0 amin0
mod amod dmod sign isign dsign int ifix idint nint idnint anint aint real float dble sngl dfloat len index ichar char
iand ior ieor ishft not btest ibset ibclr dim ddim idim dprod lge lgt lle llt len_trim trim int8 floor ceiling""".split())


class TranslationError(Exception):
    pass


# ----------------------------------------------------------------------------------------------------------------------
# program units
# ----------------------------------------------------------------------------------------------------------------------

_unit_re = re.compile(
    r"^\s*(?:(?P<typ>real\s*\*\s*\d+|integer\s*\*\s*\d+|real|integer|logical|double\s*precision|character(?:\s*\*\s*\d+)?)\s+)?"
    r"(?P<kind>subroutine|function|program|block\s*data)\s*(?P<name>\w+)?\s*(?:\((?P<args>[^)]*)\))?\s*$", re.I)


def split_units(stmts):
    """Yields (kind, name, args, rettype_text, [Stmt])."""
    units = []
    cur = None
    for st in stmts:
        low = st.text.lower()
        if cur is None:
            m = _unit_re.match(st.text)
            if not m:
                # stray statement outside a unit (e.g. leftover) -> ignore silently
                continue
            kind = re.sub(r"\s+", "", m.group("kind").lower())
            name = (m.group("name") or "blockdata").lower()
            args = [a.strip().lower() for a in (m.group("args") or "").split(",") if a.strip()]
            cur = dict(kind=kind, name=name, args=args, rettype=m.group("typ"), stmts=[], file=st.file, line=st.line)
            continue
        if re.match(r"^end\s*$", low) or re.match(r"^end\s*(subroutine|function|program)\b", low):
            units.append(cur)
            cur = None
            continue
        cur["stmts"].append(st)
    return units


def parse_type_text(t):
    """'real*8' -> ('r8',None); 'character*4' -> ('ch',4)"""
    t = re.sub(r"\s+", "", t.lower())
    if t.startswith("doubleprecision"):
        return "r8", None
    m = re.match(r"(real|integer|logical|character|complex)(?:\*(\d+|\(\*\)|\(\d+\)))?$", t)
    if not m:
        raise TranslationError(f"bad type {t!r}")
    base, k = m.group(1), m.group(2)
    if base == "real":
        return ("r4" if k == "4" else "r8"), None
    if base == "integer":
        return {None: "i4", "4": "i4", "8": "i8", "2": "i2", "1": "i1"}[k], None
    if base == "logical":
        return ("l1" if k == "1" else "l4"), None
    if base == "character":
        if k is None:
            return "ch", 1
        k = k.strip("()")
        return "ch", ("*" if k == "*" else int(k))
    raise TranslationError(f"unsupported type {t!r}")


_decl_re = re.compile(r"^(real\s*\*\s*\d+|integer\s*\*\s*\d+|logical\s*\*\s*\d+|double\s*precision|real|integer|logical|"
                      r"character\s*\*\s*\(\s*\*\s*\)|character\s*\*\s*\(\s*\w+\s*\)|character\s*\*\s*\d+|character)\s*(?![=\w(])?(.*)$", re.I)


class Unit:
    """One subroutine/function: symbol table + C code generation."""

    def __init__(self, tr, info):
        self.tr = tr
        self.kind, self.name, self.args = info["kind"], info["name"], info["args"]
        self.stmts = info["stmts"]
        self.file, self.line = info["file"], info["line"]
        self.syms = {}
        self.implicit = {}
        for c in "abcdefghopqrstuvwxyz$_":
            self.implicit[c] = ("r8", None)
        for c in "ijklmn":
            self.implicit[c] = ("i4", None)
        self.implicit_none = False
        self.commons = {}          # block -> [names in order]
        self.equivs = []           # list of lists of ASTs
        self.datas = []            # (targets ASTs, values)
        self.stfuncs = {}          # statement functions: name -> (params, ast)
        self.exec = []
        self.calls = set()
        self.used = set()
        self.tmp = 0
        self.rettype = None
        if self.kind == "function":
            s = self.sym(self.name)
            s.is_func_result = True
            if info["rettype"]:
                s.typ, s.clen = parse_type_text(info["rettype"])
                s.explicit = True
        for a in self.args:
            self.sym(a).arg = True
        self.scan()

    # ---- symbols -------------------------------------------------------------------------------------------------
    def sym(self, n) -> Sym:
        s = self.syms.get(n)
        if s is None:
            s = self.syms[n] = Sym(n)
        return s

    def typeof_sym(self, s: Sym):
        if s.typ is None:
            t, l = self.implicit.get(s.name[0], ("r8", None))
            return t
        return s.typ

    def where(self, st):
        return f"{os.path.basename(st.file)}:{st.line} ({self.name})"

    # ---- pass 1: declarations --------------------------------------------------------------------------------------
    def scan(self):
        in_decl = True
        for st in self.stmts:
            txt = st.text
            low = txt.lower()
            try:
                if self.try_decl(st, txt, low):
                    continue
            except (SyntaxError, TranslationError) as e:
                raise TranslationError(f"{self.where(st)}: {e}\n   {txt}")
            self.exec.append(st)

    def parse_entity_list(self, rest, where):
        """'a(10,2), b, c*4' -> [(name, dims|None, charlen|None)] via the expression parser."""
        toks = tokenize(rest)
        p = Parser(toks, where)
        out = []
        while not p.done():
            name = p.expect("name")[1]
            dims, clen = None, None
            if p.at("op", "("):
                p.next()
                dims = p.arglist()
            if p.at("op", "*"):
                p.next()
                if p.at("op", "("):
                    p.next()
                    if p.at("op", "*"):
                        p.next()
                        clen = "*"
                    else:
                        clen = self.const(p.expr())
                    p.expect("op", ")")
                else:
                    clen = int(p.expect("int")[1])
            out.append((name, dims, clen))
            if p.at("op", ","):
                p.next()
        return out

    def set_dims(self, s: Sym, dims):
        d = []
        for a in dims:
            if a[0] == "range":
                lo, hi = a[1], a[2]
                if hi is not None and hi[0] == "star":
                    hi = None
                d.append((lo if lo is not None else ("int", "1"), hi))
            elif a[0] == "star":
                d.append((("int", "1"), None))
            else:
                d.append((("int", "1"), a))
        s.dims = d

    def try_decl(self, st, txt, low):
        w = self.where(st)
        if low.startswith("implicit"):
            body = low[8:].strip()
            if body.replace(" ", "") == "none":
                self.implicit_none = True
                return True
            m = re.match(r"(.+?)\(([^)]*)\)\s*$", body)
            # possibly several specs: implicit real*8 (a-h,o-z), integer (i-n)
            for spec in re.findall(r"([a-z0-9* ]+?)\s*\(([a-z,\- ]+)\)", body):
                t = parse_type_text(spec[0].strip().lstrip(","))
                for rng in spec[1].replace(" ", "").split(","):
                    a, _, b = rng.partition("-")
                    b = b or a
                    for c in range(ord(a), ord(b) + 1):
                        self.implicit[chr(c)] = t
            return True
        if re.match(r"^parameter\s*\(", low):
            inner = txt[txt.index("(") + 1: txt.rindex(")")]
            p = Parser(tokenize(inner), w)
            while not p.done():
                name = p.expect("name")[1]
                p.expect("op", "=")
                e = p.expr()
                s = self.sym(name)
                v = self.const(e)
                t = self.typeof_sym(s)
                if t in ("i4", "i8", "i2", "i1") and not isinstance(v, str):
                    v = int(v)
                elif t in ("r8", "r4") and not isinstance(v, str):
                    v = float(v)
                s.param = v
                if p.at("op", ","):
                    p.next()
            return True
        if re.match(r"^common\b", low):
            body = txt[6:].strip()
            # split into /blk/ list segments
            pos = 0
            segs = []
            if not body.startswith("/"):
                body = "/_blank_/" + body
            for m in re.finditer(r"/\s*(\w*)\s*/", body):
                segs.append((m.group(1).lower() or "_blank_", m.start(), m.end()))
            for k, (blk, a, b) in enumerate(segs):
                end = segs[k + 1][1] if k + 1 < len(segs) else len(body)
                lst = body[b:end].strip().rstrip(",")
                for name, dims, clen in self.parse_entity_list(lst, w):
                    s = self.sym(name)
                    if dims is not None:
                        self.set_dims(s, dims)
                    self.commons.setdefault(blk, []).append(name)
                    s.common = (blk, None)
            return True
        if re.match(r"^dimension\b", low):
            for name, dims, clen in self.parse_entity_list(txt[9:], w):
                self.set_dims(self.sym(name), dims)
            return True
        if re.match(r"^(external|intrinsic)\b", low):
            kw = low.split()[0]
            for n in txt[len(kw):].split(","):
                n = n.strip().lower()
                if n and kw == "external":
                    self.sym(n).external = True
            return True
        if re.match(r"^save\b", low):
            body = txt[4:].strip()
            if not body:
                self.save_all = True
                for s in self.syms.values():
                    s.save = True
                self._save_all = True
            else:
                for n in body.split(","):
                    n = n.strip().lower()
                    if n.startswith("/"):
                        continue
                    self.sym(n).save = True
            return True
        if re.match(r"^equivalence\b", low):
            p = Parser(tokenize(txt[11:]), w)
            while not p.done():
                p.expect("op", "(")
                grp = p.arglist()
                self.equivs.append(grp)
                if p.at("op", ","):
                    p.next()
            return True
        if re.match(r"^data\b", low) and not re.match(r"^data\w*\s*(\(.*\))?\s*=", low):
            self.parse_data(txt[4:], w)
            return True
        m = _decl_re.match(txt)
        if m and not re.match(r"^(real|integer|logical|character)\w*\s*(\([^=]*\))?\s*=[^=]", low.replace(" ", "")):
            ttext, rest = m.group(1), m.group(2)
            if re.match(r"^\s*function\b", rest, re.I):
                return False
            if re.match(r"character\s*\*\s*\(\s*[a-z]\w*\s*\)", ttext, re.I):
                nm = re.search(r"\(\s*(\w+)\s*\)", ttext).group(1).lower()
                t, l = "ch", int(self.sym(nm).param)
            else:
                t, l = parse_type_text(ttext)
            rest = rest.strip()
            if rest.startswith(","):
                rest = rest[1:]
            for name, dims, clen in self.parse_entity_list(rest, w):
                s = self.sym(name)
                s.typ, s.explicit = t, True
                if t == "ch":
                    s.clen = clen if clen is not None else l
                if dims is not None:
                    self.set_dims(s, dims)
            return True
        return False

    def parse_data(self, body, w):
        # data a,b /1,2/, c /3*0.0/
        p = Parser(tokenize(body), w)
        while not p.done():
            targets = []
            while True:
                targets.append(p.p_prim())
                if p.at("op", ","):
                    p.next()
                    continue
                break
            p.expect("op", "/")
            vals = []
            while True:
                # value or rep*value
                neg = False
                if p.at("op", "-"):
                    p.next()
                    neg = True
                elif p.at("op", "+"):
                    p.next()
                v = p.p_prim()
                rep = 1
                if p.at("op", "*"):
                    p.next()
                    rep = int(self.const(v))
                    neg = False
                    if p.at("op", "-"):
                        p.next()
                        neg = True
                    v = p.p_prim()
                if neg:
                    v = ("un", "-", v)
                vals.extend([v] * rep)
                if p.at("op", ","):
                    p.next()
                    continue
                break
            p.expect("op", "/")
            self.datas.append((targets, vals))
            for t in targets:
                n = t[1]
                self.sym(n).data = True
                self.sym(n).save = True
            if p.at("op", ","):
                p.next()

    # ---- constant folding -------------------------------------------------------------------------------------------
    def const(self, e):
        k = e[0]
        if k == "int":
            return int(e[1])
        if k == "real":
            return float(e[1].lower().replace("d", "e"))
        if k == "str":
            return e[1]
        if k == "log":
            return bool(e[1])
        if k == "paren":
            return self.const(e[1])
        if k == "name":
            s = self.syms.get(e[1])
            if s is None or s.param is None:
                raise TranslationError(f"{e[1]} is not a constant in {self.name}")
            return s.param
        if k == "un":
            v = self.const(e[2])
            return -v if e[1] == "-" else (not v)
        if k == "bin":
            a, b = self.const(e[2]), self.const(e[3])
            op = e[1]
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "/":
                if isinstance(a, int) and isinstance(b, int):
                    q = abs(a) // abs(b)
                    return q if (a >= 0) == (b >= 0) else -q
                return a / b
            if op == "**":
                return a ** b
            if op in ("==", "!=", "<", "<=", ">", ">="):
                return eval(f"a {op} b")
            raise TranslationError(f"const op {op}")
        if k == "app" and e[1] in ("min", "max"):
            vals = [self.const(a) for a in e[2]]
            return min(vals) if e[1] == "min" else max(vals)
        if k == "app" and e[1] in ("mod",):
            a, b = [self.const(x) for x in e[2]]
            return int(a - b * int(a / b))
        raise TranslationError(f"not a constant expression: {e}")

    def is_const(self, e):
        try:
            self.const(e)
            return True
        except TranslationError:
            return False

    # ---- layout: commons / equivalence --------------------------------------------------------------------------------
    def elem_size(self, s: Sym):
        t = self.typeof_sym(s)
        if t == "ch":
            if s.clen == "*":
                return 1
            return int(s.clen or 1)
        return TSIZE[t]

    def const_extents(self, s: Sym):
        ext = []
        for lo, hi in s.dims:
            if hi is None:
                raise TranslationError(f"assumed-size array {s.name} needs storage in {self.name}")
            ext.append(self.const(hi) - self.const(lo) + 1)
        return ext

    def storage_bytes(self, s: Sym):
        n = 1
        if s.dims:
            for x in self.const_extents(s):
                n *= x
        return n * self.elem_size(s)

    def layout(self):
        for blk, names in self.commons.items():
            off = 0
            for n in names:
                s = self.syms[n]
                s.common = (blk, off)
                off += self.storage_bytes(s)
            self.tr.common_size[blk] = max(self.tr.common_size.get(blk, 0), off)
            self.tr.common_views.setdefault(blk, {})[self.name] = [
                (n, self.syms[n].common[1], self.typeof_sym(self.syms[n]), self.elem_size(self.syms[n]),
                 self.const_extents(self.syms[n]) if self.syms[n].dims else [],
                 [self.const(lo) for lo, _ in self.syms[n].dims] if self.syms[n].dims else []) for n in names]
        # equivalence: resolve to (anchor, byte offset); anchors in COMMON give every member a COMMON address
        k = 0
        for grp in self.equivs:
            members = []
            for e in grp:
                if e[0] == "name":
                    members.append((self.sym(e[1]), 0))
                elif e[0] == "app":
                    s = self.sym(e[1])
                    idx = self.const_index(s, e[2])
                    members.append((s, idx * self.elem_size(s)))
                else:
                    raise TranslationError(f"equivalence member {e}")
            anchor = next(((s, o) for s, o in members if s.common and s.common[1] is not None), None)
            if anchor is None:
                anchor = next(((s, o) for s, o in members if s.eqv), None)
            if anchor is not None and anchor[0].common and anchor[0].common[1] is not None:
                blk, base = anchor[0].common
                for s, o in members:
                    if s is anchor[0]:
                        continue
                    addr = base + anchor[1] - o
                    if addr < 0:
                        raise TranslationError("equivalence extends a COMMON block backwards")
                    s.common = (blk, addr)
                    self.tr.common_size[blk] = max(self.tr.common_size[blk], addr + self.storage_bytes(s))
            elif anchor is not None:
                blob, base = anchor[0].eqv
                for s, o in members:
                    if s is anchor[0]:
                        continue
                    s.eqv = (blob, base + anchor[1] - o)
                    self.eqv_blobs[blob] = max(self.eqv_blobs[blob], s.eqv[1] + self.storage_bytes(s))
            else:
                blob = f"eqv{k}"
                k += 1
                shift = max(o for _, o in members)
                size = 0
                for s, o in members:
                    s.eqv = (blob, shift - o)
                    size = max(size, s.eqv[1] + self.storage_bytes(s))
                self.eqv_blobs[blob] = size

    def const_index(self, s: Sym, subs):
        ext = self.const_extents(s) if all(h is not None for _, h in s.dims) else None
        idx, mul = 0, 1
        for d, a in enumerate(subs):
            lo = self.const(s.dims[d][0])
            idx += (self.const(a) - lo) * mul
            if d + 1 < len(subs):
                mul *= ext[d]
        return idx

    # ---- expression typing + C generation -------------------------------------------------------------------------------
    def new_tmp(self, p="t"):
        self.tmp += 1
        return f"_{p}{self.tmp}"

    def is_ptr(self, s: Sym):
        """True when the C identifier of `s` is a pointer (dummy argument, COMMON or EQUIVALENCE member, array)."""
        return s.arg or s.common is not None or s.eqv is not None or s.is_array or self.typeof_sym(s) == "ch"

    def typeof(self, e):
        k = e[0]
        if k == "int":
            return "i4" if abs(int(e[1])) < 2 ** 31 else "i8"
        if k == "real":
            return "r8"
        if k == "str":
            return "ch"
        if k == "log":
            return "l4"
        if k == "paren":
            return self.typeof(e[1])
        if k == "name":
            return self.typeof_sym(self.sym(e[1]))
        if k == "substr":
            return "ch"
        if k == "un":
            return "l4" if e[1] == "!" else self.typeof(e[2])
        if k == "bin":
            op = e[1]
            if op in ("==", "!=", "<", "<=", ">", ">=", "&&", "||", "eqv", "neqv"):
                return "l4"
            if op == "//":
                return "ch"
            ta, tb = self.typeof(e[2]), self.typeof(e[3])
            if op == "**" and RANK.get(tb, 1) <= 2:
                return ta if RANK.get(ta, 1) >= 1 else "i4"
            return ta if RANK.get(ta, 0) >= RANK.get(tb, 0) else tb
        if k == "app":
            n = e[1]
            s = self.syms.get(n)
            if s is not None and (s.is_array or (self.typeof_sym(s) == "ch" and not s.external and e[2] and e[2][0][0] == "range")):
                return self.typeof_sym(s)
            if n in self.stfuncs:
                return self.typeof_sym(self.sym(n))
            if n in INTRINSICS and not (s and (s.external or s.is_array)):
                return self.intrinsic_type(n, e[2])
            return self.typeof_sym(self.sym(n))
        raise TranslationError(f"typeof {e}")

    def intrinsic_type(self, n, args):
        if n in ("int", "ifix", "idint", "nint", "idnint", "iabs", "max0", "min0", "isign", "len", "index", "ichar", "iand",
                 "ior", "ieor", "ishft", "not", "ibset", "ibclr", "idim", "len_trim", "floor", "ceiling"):
            if n in ("iand", "ior", "ieor", "ishft", "not", "ibset", "ibclr") and args:
                return self.typeof(args[0])
            return "i4"
        if n == "int8":
            return "i8"
        if n in ("btest", "lge", "lgt", "lle", "llt"):
            return "l4"
        if n == "char":
            return "ch"
        if n in ("abs", "max", "min", "mod", "sign", "dim"):
            t = "l4"
            for a in args:
                ta = self.typeof(a)
                if RANK.get(ta, 0) > RANK.get(t, 0):
                    t = ta
            return t
        return "r8"

    def cx(self, e):
        """C text of an expression (rvalue)."""
        k = e[0]
        if k == "int":
            v = int(e[1])
            return str(v) if abs(v) < 2 ** 31 else f"{v}LL"
        if k == "real":
            t = e[1].lower().replace("d", "e")
            if "e" not in t and "." not in t:
                t += ".0"
            return t
        if k == "log":
            return str(e[1])
        if k == "str":
            raise TranslationError("character literal in numeric context")
        if k == "paren":
            return "(" + self.cx(e[1]) + ")"
        if k == "name":
            s = self.sym(e[1])
            self.used.add(e[1])
            if s.param is not None:
                return self.cconst(s.param)
            if s.is_array:
                raise TranslationError(f"whole-array reference {e[1]} in expression")
            if s.is_func_result:
                return f"{cname(e[1])}_result"
            return f"(*{cname(e[1])})" if self.is_ptr(s) else cname(e[1])
        if k == "un":
            if e[1] == "-":
                return "(-" + self.cx(e[2]) + ")"
            return "(!" + self.cx(e[2]) + ")"
        if k == "bin":
            return self.cbin(e)
        if k == "app":
            return self.capp(e)
        raise TranslationError(f"cannot translate {e}")

    def cconst(self, v):
        if isinstance(v, bool):
            return "1" if v else "0"
        if isinstance(v, int):
            return str(v) if abs(v) < 2 ** 31 else f"{v}LL"
        if isinstance(v, float):
            return repr(v)
        raise TranslationError(f"constant {v!r} in numeric context")

    def cbin(self, e):
        op, a, b = e[1], e[2], e[3]
        ta, tb = self.typeof(a), self.typeof(b)
        if ta == "ch" or tb == "ch":
            if op in ("==", "!="):
                pa, la = self.cstr(a)
                pb, lb = self.cstr(b)
                r = f"f77_streq({pa},{la},{pb},{lb})"
                return r if op == "==" else f"(!{r})"
            raise TranslationError(f"character operator {op} unsupported")
        if op == "**":
            if RANK.get(tb, 1) <= 2:
                if RANK.get(ta, 1) <= 2:
                    return f"f77_ipow({self.cx(a)},{self.cx(b)})"
                if self.is_const(b) and self.const(b) == 2:
                    ca = self.cx(a)
                    if a[0] in ("name", "app", "paren", "real"):
                        # x**2 == x*x exactly what gfortran emits
                        if a[0] == "app" and not self.is_arrayref(a):
                            return f"__builtin_powi({ca},2)"
                        return f"({ca}*{ca})"
                return f"__builtin_powi({self.cx(a)},{self.cx(b)})"
            return f"pow({self.cx(a)},{self.cx(b)})"
        if op in ("eqv", "neqv"):
            return f"((!!{self.cx(a)}) {'==' if op == 'eqv' else '!='} (!!{self.cx(b)}))"
        ca, cb_ = self.cx(a), self.cx(b)
        if op == "/" and RANK.get(ta, 1) <= 2 and RANK.get(tb, 1) <= 2:
            return f"({ca}/{cb_})"      # C integer division truncates toward zero like Fortran
        return f"({ca} {op} {cb_})"

    def is_arrayref(self, e):
        if e[0] != "app":
            return False
        s = self.syms.get(e[1])
        return s is not None and s.is_array

    def index_c(self, s: Sym, subs, w=""):
        """Zero-based linear index (C text) of array element."""
        if len(subs) > len(s.dims):
            raise TranslationError(f"{s.name}: {len(subs)} subscripts for rank {len(s.dims)}")
        terms = None
        # Horner from the last subscript
        n = len(subs)
        for d in range(n - 1, -1, -1):
            lo = s.dims[d][0]
            sub = self.cx(subs[d])
            if self.is_const(lo):
                lov = self.const(lo)
                t = f"({sub})" if lov == 0 else f"({sub}-{lov})" if lov > 0 else f"({sub}+{-lov})"
            else:
                t = f"({sub}-({self.cx(lo)}))"
            if terms is None:
                terms = f"(long){t}"
            else:
                terms = f"{t}+{self.extent_c(s, d)}*({terms})"
        return terms

    def extent_c(self, s: Sym, d):
        lo, hi = s.dims[d]
        if hi is None:
            raise TranslationError(f"{s.name}: extent of assumed-size dimension needed")
        if self.is_const(lo) and self.is_const(hi):
            return str(self.const(hi) - self.const(lo) + 1)
        return f"{cname(s.name)}_x{d}"

    def capp(self, e):
        n, args = e[1], e[2]
        s = self.syms.get(n)
        if s is not None and s.is_array:
            self.used.add(n)
            if self.typeof_sym(s) == "ch":
                raise TranslationError("character array element in numeric context")
            return f"{cname(n)}[{self.index_c(s, args)}]"
        if n in self.stfuncs:
            params, body = self.stfuncs[n]
            return "(" + self.cx(subst(body, dict(zip(params, args)))) + ")"
        if n in INTRINSICS and not (s and s.external):
            return self.cintrinsic(n, args)
        # external function
        return self.ccall(n, args, is_func=True)

    def cintrinsic(self, n, args):
        ts = [self.typeof(a) for a in args]
        cs = [self.cx(a) if t != "ch" else None for a, t in zip(args, ts)]
        isint = all(RANK.get(t, 1) <= 2 for t in ts)
        if n in ("abs", "iabs", "dabs"):
            return f"fabs({cs[0]})" if not isint else (f"llabs({cs[0]})" if ts[0] == "i8" else f"abs({cs[0]})")
        simple = {"sqrt": "sqrt", "dsqrt": "sqrt", "exp": "exp", "dexp": "exp", "log": "log", "alog": "log", "dlog": "log",
                  "log10": "log10", "alog10": "log10", "dlog10": "log10", "sin": "sin", "dsin": "sin", "cos": "cos",
                  "dcos": "cos", "tan": "tan", "dtan": "tan", "asin": "asin", "dasin": "asin", "acos": "acos",
                  "dacos": "acos", "atan": "atan", "datan": "atan", "atan2": "atan2", "datan2": "atan2", "sinh": "sinh",
                  "dsinh": "sinh", "cosh": "cosh", "dcosh": "cosh", "tanh": "tanh", "dtanh": "tanh", "anint": "round",
                  "aint": "trunc"}
        if n in simple:
            return f"{simple[n]}({','.join('(double)' + c for c in cs)})"
        if n in ("max", "min", "amax1", "amin1", "max0", "min0", "dmax1", "dmin1", "amax0", "amin0"):
            big = "max" in n
            f = ("f77_dmax" if big else "f77_dmin") if not isint else (
                ("f77_lmax" if big else "f77_lmin") if "i8" in ts else ("f77_imax" if big else "f77_imin"))
            r = cs[0]
            for c in cs[1:]:
                r = f"{f}({r},{c})"
            return r
        if n in ("mod", "amod", "dmod"):
            return f"({cs[0]} % {cs[1]})" if isint else f"fmod({cs[0]},{cs[1]})"
        if n in ("sign", "isign", "dsign"):
            return f"f77_isign({cs[0]},{cs[1]})" if isint else f"f77_dsign({cs[0]},{cs[1]})"
        if n in ("dim", "ddim", "idim"):
            return f"f77_imax({cs[0]}-{cs[1]},0)" if isint else f"f77_dmax({cs[0]}-{cs[1]},0.0)"
        if n in ("int", "ifix", "idint"):
            return f"((int)({cs[0]}))"
        if n == "int8":
            return f"((long long)({cs[0]}))"
        if n in ("nint", "idnint"):
            return f"((int)lround({cs[0]}))"
        if n == "floor":
            return f"((int)floor({cs[0]}))"
        if n == "ceiling":
            return f"((int)ceil({cs[0]}))"
        if n in ("real", "float", "dble", "sngl", "dfloat", "dprod"):
            if n == "dprod":
                return f"((double)({cs[0]})*(double)({cs[1]}))"
            return f"((double)({cs[0]}))"
        if n in ("len", "len_trim"):
            p, l = self.cstr(args[0])
            return f"({l})" if n == "len" else f"f77_len_trim({p},{l})"
        if n == "ichar":
            p, l = self.cstr(args[0])
            return f"((int)(unsigned char)({p})[0])"
        if n == "index":
            pa, la = self.cstr(args[0])
            pb, lb = self.cstr(args[1])
            return f"f77_index({pa},{la},{pb},{lb})"
        if n == "iand":
            return f"({cs[0]} & {cs[1]})"
        if n == "ior":
            return f"({cs[0]} | {cs[1]})"
        if n == "ieor":
            return f"({cs[0]} ^ {cs[1]})"
        if n == "not":
            return f"(~{cs[0]})"
        if n == "ishft":
            return f"f77_ishft({cs[0]},{cs[1]})"
        if n == "btest":
            return f"((({cs[0]}) >> ({cs[1]})) & 1)"
        if n == "ibset":
            return f"(({cs[0]}) | (1 << ({cs[1]})))"
        if n == "ibclr":
            return f"(({cs[0]}) & ~(1 << ({cs[1]})))"
        raise TranslationError(f"intrinsic {n} unsupported")

    # character values -> (pointer text, length text)
    def cstr(self, e):
        k = e[0]
        if k == "str":
            return '"' + e[1].replace("\\", "\\\\").replace('"', '\\"') + '"', str(len(e[1]))
        if k == "paren":
            return self.cstr(e[1])
        if k == "name":
            s = self.sym(e[1])
            self.used.add(e[1])
            if s.param is not None:
                return self.cstr(("str", s.param))
            return cname(e[1]), self.clen_c(s)
        if k == "app":
            s = self.syms.get(e[1])
            if s is not None and s.is_array:
                self.used.add(e[1])
                l = self.clen_c(s)
                return f"({cname(e[1])}+({l})*({self.index_c(s, e[2])}))", l
            if s is not None and self.typeof_sym(s) == "ch" and e[2] and e[2][0][0] == "range":
                self.used.add(e[1])
                return self.csubstr(cname(e[1]), self.clen_c(s), e[2][0])
            if e[1] == "char":
                return f"(char[1]){{(char)({self.cx(e[2][0])})}}", "1"
            raise TranslationError(f"character function {e[1]} unsupported")
        if k == "substr":
            p, l = self.cstr(e[1])
            return self.csubstr(p, l, e[2])
        raise TranslationError(f"character expression {e} unsupported")

    def csubstr(self, p, l, rng):
        lo = self.cx(rng[1]) if rng[1] is not None else "1"
        hi = self.cx(rng[2]) if rng[2] is not None else l
        return f"({p}+({lo})-1)", f"(({hi})-({lo})+1)"

    def clen_c(self, s: Sym):
        if s.clen == "*":
            return f"{cname(s.name)}_len"
        return str(int(s.clen or 1))

    # ---- calls ------------------------------------------------------------------------------------------------------
    def carg(self, a, hidden):
        """C text passing `a` by reference; appends hidden character lengths to `hidden`."""
        k = a[0]
        if k == "name":
            s = self.sym(a[1])
            self.used.add(a[1])
            if s.param is not None:
                if isinstance(s.param, str):
                    p, l = self.cstr(("str", s.param))
                    hidden.append(l)
                    return p
                t = self.typeof_sym(s)
                return f"&({CTYPE[t]}){{{self.cconst(s.param)}}}"
            if s.external or (a[1] in self.tr.known_units and not s.explicit and not s.arg and s.common is None and
                              a[1] not in self.assigned and not s.is_array):
                # procedure passed as argument
                self.calls.add(a[1])
                self.tr.note_extern(a[1], "void" if self.tr.unit_kind(a[1]) == "subroutine" else None, self)
                return f"(void*){a[1]}_"
            if self.typeof_sym(s) == "ch":
                hidden.append(self.clen_c(s))
                return cname(a[1])
            if s.is_func_result:
                return f"&{cname(a[1])}_result"
            return cname(a[1]) if self.is_ptr(s) else f"&{cname(a[1])}"
        if k == "app":
            s = self.syms.get(a[1])
            if s is not None and s.is_array:
                self.used.add(a[1])
                if self.typeof_sym(s) == "ch":
                    p, l = self.cstr(a)
                    hidden.append(l)
                    return p
                return f"&{cname(a[1])}[{self.index_c(s, a[2])}]"
            if s is not None and self.typeof_sym(s) == "ch" and a[2] and a[2][0][0] == "range":
                p, l = self.cstr(a)
                hidden.append(l)
                return p
        if k in ("str", "substr") or self.typeof(a) == "ch":
            p, l = self.cstr(a)
            hidden.append(l)
            return p
        t = self.typeof(a)
        return f"&({CTYPE[t]}){{{self.cx(a)}}}"

    def ccall(self, n, args, is_func=False):
        hidden = []
        cargs = [self.carg(a, hidden) for a in args]
        self.calls.add(n)
        s = self.sym(n)
        if s.arg:
            # dummy procedure
            rt = CTYPE[self.typeof_sym(s)] if is_func else "void"
            return f"(({rt}(*)())({cname(n)}))({','.join(cargs + ['(long)' + h for h in hidden])})"
        rt = CTYPE[self.typeof_sym(s)] if is_func else "void"
        self.tr.note_extern(n, rt, self)
        return f"{n}_({','.join(cargs + ['(long)' + h for h in hidden])})"

    # ---- statements ---------------------------------------------------------------------------------------------------
    def lvalue(self, e):
        k = e[0]
        if k == "name":
            s = self.sym(e[1])
            self.used.add(e[1])
            if s.is_func_result:
                return f"{cname(e[1])}_result"
            if s.param is not None:
                raise TranslationError(f"assignment to parameter {e[1]}")
            return f"(*{cname(e[1])})" if self.is_ptr(s) else cname(e[1])
        if k == "app":
            s = self.syms.get(e[1])
            if s is None or not s.is_array:
                raise TranslationError(f"assignment to non-array {e[1]}(...) (statement function?)")
            self.used.add(e[1])
            return f"{cname(e[1])}[{self.index_c(s, e[2])}]"
        raise TranslationError(f"bad lvalue {e}")

    def gen_assign(self, lhs, rhs):
        tl = self.typeof(lhs)
        if tl == "ch":
            pd, ld = self.cstr(lhs)
            if rhs[0] == "bin" and rhs[1] == "//":
                raise TranslationError("character concatenation unsupported")
            ps, ls = self.cstr(rhs)
            return f"f77_strcpy({pd},{ld},{ps},{ls});"
        tr_ = self.typeof(rhs)
        c = self.cx(rhs)
        if RANK.get(tl, 1) <= 2 and RANK.get(tr_, 1) >= 3:
            c = f"({CTYPE[tl]})({c})"
        return f"{self.lvalue(lhs)} = {c};"

    def find_assigned(self):
        """Names assigned or used as DO variables (to tell scalars from procedure names passed as arguments)."""
        self.assigned = set()
        for st in self.exec:
            m = re.match(r"^(?:do\s*\d*\s*,?\s*)?([a-z_$][\w$]*)\s*(\(.*\))?\s*=", st.text, re.I)
            if m:
                self.assigned.add(m.group(1).lower())

    def translate(self):
        """Returns the C text of the unit."""
        self.eqv_blobs = {}
        self.layout()
        self.find_assigned()
        body = []
        stack = []   # ('do', label) | ('if',)
        ind = 1

        def emit(s):
            body.append("  " * (ind) + s)

        labels_used = set()
        for st in self.exec:
            for m in re.finditer(r"go\s*to\s*(\d+)", st.text, re.I):
                labels_used.add(m.group(1))
        i = 0
        nst = len(self.exec)
        while i < nst:
            st = self.exec[i]
            i += 1
            w = self.where(st)
            try:
                txt = st.text
                low = txt.lower()
                lab = st.label
                if lab and re.match(r"^format\s*\(", low):
                    continue
                if lab:
                    lab = str(int(lab))
                    emit(f"L{lab}:;")
                c = self.gen_stmt(txt, low, stack, w)
                for line in c:
                    if line.startswith("}"):
                        ind -= line.count("}") - line.count("{") if line.count("}") > line.count("{") else 0
                    if line.startswith("} else"):
                        ind -= 0
                    emit(line)
                    if line.endswith("{"):
                        ind += 1 if not line.startswith("}") else 1
                    elif line.count("{") > line.count("}"):
                        ind += line.count("{") - line.count("}")
                # close DO loops ending on this label
                while lab and stack and stack[-1][0] == "do" and stack[-1][1] == lab:
                    stack.pop()
                    ind -= 1
                    emit("}}")
            except (SyntaxError, TranslationError) as e:
                raise TranslationError(f"{w}: {e}\n   {st.text}")
        if stack:
            raise TranslationError(f"{self.name}: unterminated block {stack}")
        return self.prologue() + body + self.epilogue()

    def gen_stmt(self, txt, low, stack, w):
        nosp = low.replace(" ", "")
        # ---- block structure
        if re.match(r"^end\s*do\b", low):
            if not stack or stack[-1][0] != "do":
                raise TranslationError("enddo without do")
            stack.pop()
            return ["}}"]
        if re.match(r"^end\s*if\b", low):
            if not stack or stack[-1][0] != "if":
                raise TranslationError("endif without if")
            stack.pop()
            return ["}"]
        if re.match(r"^else\s*if\s*\(", low) and nosp.endswith(")then"):
            cond = txt[txt.index("("): txt.lower().rindex("then")].strip()
            e = self.parse_paren_expr(cond, w)
            return [f"}} else if ({self.cx(e)}) {{"]
        if nosp == "else":
            return ["} else {"]
        if re.match(r"^if\s*\(", low):
            close = match_paren(txt, txt.index("("))
            cond = txt[txt.index("(") + 1: close]
            rest = txt[close + 1:].strip()
            e = Parser(tokenize(cond), w)
            ce = e.expr()
            if not e.done():
                raise SyntaxError(f"trailing tokens in condition {cond}")
            if rest.lower().replace(" ", "") == "then":
                stack.append(("if",))
                return [f"if ({self.cx(ce)}) {{"]
            if re.match(r"^\d+\s*,\s*\d+\s*,\s*\d+$", rest):
                l1, l2, l3 = [str(int(x)) for x in rest.split(",")]
                t = self.new_tmp()
                return [f"{{ double {t} = {self.cx(ce)}; if ({t} < 0) goto L{l1}; else if ({t} == 0) goto L{l2}; else goto L{l3}; }}"]
            inner = self.gen_stmt(rest, rest.lower(), stack, w)
            return [f"if ({self.cx(ce)}) {{"] + inner + ["}"]
        m = re.match(r"^do\s*(\d+)?\s*,?\s*([a-z_$][\w$]*)\s*=(.*)$", txt, re.I)
        if m and self.has_top_comma(m.group(3)):
            lab, var, rest = m.group(1), m.group(2).lower(), m.group(3)
            p = Parser(tokenize(rest), w)
            lo = p.expr()
            p.expect("op", ",")
            hi = p.expr()
            step = None
            if p.at("op", ","):
                p.next()
                step = p.expr()
            v = self.lvalue(("name", var))
            tv = CTYPE[self.typeof(("name", var))]
            te = self.new_tmp("e")
            stack.append(("do", str(int(lab)) if lab else None))
            if step is None:
                return [f"{{ {tv} {te} = {self.cx(hi)}; for ({v} = {self.cx(lo)}; {v} <= {te}; {v}++) {{"]
            ts = self.new_tmp("s")
            return [f"{{ {tv} {te} = {self.cx(hi)}; {tv} {ts} = {self.cx(step)}; "
                    f"for ({v} = {self.cx(lo)}; ({ts} > 0) ? ({v} <= {te}) : ({v} >= {te}); {v} += {ts}) {{"]
        m = re.match(r"^do\s*while\s*\((.*)\)\s*$", txt, re.I)
        if m:
            e = Parser(tokenize(m.group(1)), w).expr()
            stack.append(("do", None))
            return [f"{{ while ({self.cx(e)}) {{"]
        # ---- simple statements
        if nosp == "continue":
            return [";"]
        if nosp == "return":
            return [self.c_return()]
        if nosp == "exit":
            return ["break;"]
        if nosp == "cycle":
            return ["continue;"]
        if re.match(r"^stop\b", low):
            return ["f77_stop();"]
        m = re.match(r"^go\s*to\s*(\d+)\s*$", low)
        if m:
            return [f"goto L{int(m.group(1))};"]
        m = re.match(r"^go\s*to\s*\(([\d,\s]+)\)\s*,?\s*(.+)$", low)
        if m:
            labs = [str(int(x)) for x in m.group(1).split(",")]
            e = Parser(tokenize(m.group(2)), w).expr()
            t = self.new_tmp()
            return [f"{{ int {t} = {self.cx(e)}; " + " ".join(f"if ({t} == {k + 1}) goto L{l};" for k, l in enumerate(labs)) + " }"]
        if re.match(r"^(write|read)\s*\(", low) or re.match(r"^print\b", low) or \
           re.match(r"^(open|close|rewind|backspace|endfile|inquire|flush)\s*\(", low) or re.match(r"^format\s*\(", low):
            if low.startswith("read"):
                return ['f77_unsupported("read statement");']
            tr = self.trace_write(txt, low)
            return [tr] if tr else ["/* i/o dropped */;"]
        m = re.match(r"^call\s+([a-z_$][\w$]*)\s*(\(.*\))?\s*$", txt, re.I)
        if m:
            name = m.group(1).lower()
            args = []
            if m.group(2):
                p = Parser(tokenize(m.group(2)), w)
                p.expect("op", "(")
                args = p.arglist()
                if not p.done():
                    raise SyntaxError("trailing tokens after call")
            return [self.ccall(name, args) + ";"]
        # assignment (or statement function definition)
        p = Parser(tokenize(txt), w)
        lhs = p.p_prim()
        if not p.at("op", "="):
            raise TranslationError(f"unrecognised statement")
        p.next()
        rhs = p.expr()
        if not p.done():
            raise SyntaxError("trailing tokens after assignment")
        if lhs[0] == "app" and not self.is_arrayref(lhs) and not (self.typeof_sym(self.sym(lhs[1])) == "ch"):
            # statement function definition
            self.stfuncs[lhs[1]] = ([a[1] for a in lhs[2]], rhs)
            return [f"/* statement function {lhs[1]} */;"]
        return [self.gen_assign(lhs, rhs)]

    def trace_write(self, txt, low):
        """write(6,...) / write(*,...) in a routine named in Translator.trace_units: instead of being dropped, the NUMERIC scalar
        items of the output list are handed to f77_trace(unit, n, values...) (oracle/ref_stubs.c keeps them in a ring buffer
        when tracing is switched on).  This is how the tests read what the reference LOGS -- e.g. cggo's per-iteration residual
        (core/hmholtz.f:770-773), which it keeps in no COMMON variable.  Strings, implied-do lists and anything that does not
        parse are skipped; the arithmetic of the routine is untouched."""
        if self.name not in getattr(self.tr, "trace_units", ()) or not low.startswith("write"):
            return None
        i = low.index("(")
        j = match_paren(txt, i)
        if j < 0:
            return None
        ctl = re.sub(r"\s+", "", low[i + 1:j])
        if not (ctl.startswith("6,") or ctl.startswith("*,") or ctl in ("6", "*")):
            return None
        rest = txt[j + 1:].strip()
        if not rest:
            return None
        try:
            p = Parser(tokenize("(" + rest + ")"), "")
            p.expect("op", "(")
            items = p.arglist()
            if not p.done():
                return None
            vals = []
            for e in items:
                if e[0] in ("range", "star"):
                    return None
                t = self.typeof(e)
                if t == "ch":
                    continue
                if e[0] == "app" and not self.is_arrayref(e) and e[1] not in self.stfuncs and e[1] not in INTRINSICS:
                    return None                     # a function call in an output list: leave it alone
                if e[0] == "name" and self.sym(e[1]).is_array:
                    return None                     # whole-array output
                vals.append(f"(double)({self.cx(e)})")
            if not vals or len(vals) > 12:
                return None
            return f'if (f77_trace_on) f77_trace("{self.name}", {len(vals)}, {", ".join(vals)});'
        except (TranslationError, SyntaxError, KeyError, IndexError):
            return None

    def has_top_comma(self, s):
        depth, q = 0, None
        for ch in s:
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 0:
                return True
        return False

    def parse_paren_expr(self, s, w):
        p = Parser(tokenize(s), w)
        e = p.expr()
        if not p.done():
            raise SyntaxError(f"trailing tokens in {s}")
        return e

    def c_return(self):
        fin = " ".join(self.frees) if getattr(self, "frees", None) else ""
        if self.kind == "function":
            return f"{fin} return {cname(self.name)}_result;"
        return f"{fin} return;"

    # ---- prologue: signature, local declarations, views --------------------------------------------------------------------
    def prologue(self):
        out = []
        self.frees = []
        hidden = [a for a in self.args if self.typeof_sym(self.syms[a]) == "ch" and not self.syms[a].external]
        params = [f"void *{cname(a)}_a" for a in self.args] + [f"long {cname(a)}_len" for a in hidden]
        rt = "void"
        if self.kind == "function":
            rt = CTYPE[self.typeof_sym(self.syms[self.name])]
        self.tr.defined[self.name] = rt
        out.append(f"/* {os.path.relpath(self.file, self.tr.root)}:{self.line} */")
        out.append(f"{rt} {self.name}_({', '.join(params) if params else 'void'})")
        out.append("{")
        decl = []
        if self.kind == "function":
            decl.append(f"  {rt} {cname(self.name)}_result = 0;")
        for blob, size in self.eqv_blobs.items():
            decl.append(f"  static char {blob}[{size}] __attribute__((aligned(16)));")
        # DATA for variables the body never touches (include-file boilerplate) is dropped
        live = []
        for targets, vals in self.datas:
            if any(t[1] in self.used for t in targets):
                live.append((targets, vals))
                for t in targets:
                    self.used.add(t[1])
        self.datas = live
        # names that only appear in the bounds of used arrays must be declared too
        changed = True
        while changed:
            before = len(self.used)
            for n in list(self.used) + list(self.args):
                sy = self.syms.get(n)
                if sy is not None and sy.dims:
                    for lo, hi in sy.dims:
                        for b in (lo, hi):
                            if b is not None and not self.is_const(b):
                                self.cx(b)
            changed = len(self.used) != before
        # scalars first (array extents may need them), then arrays
        names = [n for n in self.syms if n in self.used or self.syms[n].arg]
        late = []
        for n in names:
            s = self.syms[n]
            if s.param is not None or s.is_func_result:
                continue
            t = self.typeof_sym(s)
            ct = CTYPE[t]
            c = cname(n)
            if s.external or (n in self.calls and not s.is_array and not s.arg):
                continue
            if s.arg and (s.external or (n in self.calls and not s.is_array)):
                decl.append(f"  void *{c} = {c}_a;")
                continue
            if n in INTRINSICS and not s.explicit and not s.arg and not s.common and not s.is_array and n not in self.assigned:
                continue
            if s.arg:
                decl.append(f"  {ct} *{c} = ({ct}*){c}_a;")
            elif s.common is not None:
                if s.common[1] is None:
                    raise TranslationError(f"{self.name}: COMMON offset of {n} unresolved")
                self.tr.common_used.add(s.common[0])
                decl.append(f"  {ct} *{c} = ({ct}*)(cb_{s.common[0]} + {s.common[1]});")
            elif s.eqv is not None:
                decl.append(f"  {ct} *{c} = ({ct}*)({s.eqv[0]} + {s.eqv[1]});")
            elif s.is_array:
                if all(self.is_const(lo) and hi is not None and self.is_const(hi) for lo, hi in s.dims):
                    tot = 1
                    for x in self.const_extents(s):
                        tot *= x
                    tot *= (int(s.clen) if t == "ch" else 1)
                    decl.append(f"  static {ct} {c}[{max(tot, 1)}];")
                else:
                    # automatic array: extents depend on arguments
                    late.append(("auto", s))
                    continue
            elif t == "ch":
                decl.append(f"  {'static ' if s.save else ''}char {c}[{int(s.clen or 1)}];")
            else:
                decl.append(f"  {'static ' if s.save else ''}{ct} {c} = 0;")
            if s.is_array and (s.arg or s.common is not None or s.eqv is not None):
                late.append(("ext", s))
        out += decl
        for kind, s in late:
            c = cname(s.name)
            for d, (lo, hi) in enumerate(s.dims):
                if hi is None:
                    continue
                if not (self.is_const(lo) and self.is_const(hi)):
                    if d + 1 < len(s.dims) or kind == "auto":
                        out.append(f"  const long {c}_x{d} = (long)({self.cx(hi)}) - (long)({self.cx(lo)}) + 1;")
            if kind == "auto":
                t = self.typeof_sym(s)
                tot = "*".join(self.extent_c(s, d) for d in range(len(s.dims)))
                out.append(f"  {CTYPE[t]} *{c} = ({CTYPE[t]}*)calloc((size_t)f77_lmax({tot},1), sizeof({CTYPE[t]}));")
                self.frees.append(f"free({c});")
        # DATA initialisation, once
        if self.datas:
            out.append("  static int f77_first = 1;")
            out.append("  if (f77_first) { f77_first = 0;")
            for targets, vals in self.datas:
                vi = 0
                for t in targets:
                    s = self.sym(t[1])
                    if t[0] == "name" and s.is_array:
                        tot = 1
                        for x in self.const_extents(s):
                            tot *= x
                        if self.typeof_sym(s) == "ch":
                            l = int(s.clen)
                            for k in range(tot):
                                ps, ls = self.cstr(vals[vi])
                                out.append(f"    f77_strcpy({cname(t[1])}+{k * l},{l},{ps},{ls});")
                                vi += 1
                        else:
                            for k in range(tot):
                                out.append(f"    {cname(t[1])}[{k}] = {self.cx(vals[vi])};")
                                vi += 1
                    else:
                        out.append("    " + self.gen_assign(t, vals[vi]))
                        vi += 1
            out.append("  }")
        return out

    def epilogue(self):
        return ["  " + self.c_return(), "}", ""]


def subst(e, env):
    if e[0] == "name" and e[1] in env:
        return ("paren", env[e[1]])
    if e[0] in ("bin",):
        return ("bin", e[1], subst(e[2], env), subst(e[3], env))
    if e[0] == "un":
        return ("un", e[1], subst(e[2], env))
    if e[0] == "paren":
        return ("paren", subst(e[1], env))
    if e[0] == "app":
        return ("app", e[1], [subst(a, env) for a in e[2]])
    return e


def match_paren(s, i):
    depth, q = 0, None
    for j in range(i, len(s)):
        ch = s[j]
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return j
    raise SyntaxError(f"unbalanced parentheses in {s!r}")


# ----------------------------------------------------------------------------------------------------------------------
# driver
# ----------------------------------------------------------------------------------------------------------------------

RUNTIME = r"""
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#undef linux
#undef unix
static inline int f77_imax(int a, int b) { return a > b ? a : b; }
static inline int f77_imin(int a, int b) { return a < b ? a : b; }
static inline long long f77_lmax(long long a, long long b) { return a > b ? a : b; }
static inline long long f77_lmin(long long a, long long b) { return a < b ? a : b; }
static inline double f77_dmax(double a, double b) { return a > b ? a : b; }
static inline double f77_dmin(double a, double b) { return a < b ? a : b; }
static inline long long f77_isign(long long a, long long b) { long long x = a < 0 ? -a : a; return b >= 0 ? x : -x; }
static inline double f77_dsign(double a, double b) { return copysign(fabs(a), b); }
static inline long long f77_ipow(long long a, long long b) { long long r = 1; if (b < 0) return (a == 1) ? 1 : (a == -1 ? ((b & 1) ? -1 : 1) : 0); while (b-- > 0) r *= a; return r; }
static inline int f77_ishft(int a, int s) { return s >= 0 ? (int)((unsigned)a << s) : (int)((unsigned)a >> (-s)); }
static int f77_streq(const char *a, long la, const char *b, long lb) {
  long n = la < lb ? la : lb, i;
  for (i = 0; i < n; i++) if (a[i] != b[i]) return 0;
  for (i = n; i < la; i++) if (a[i] != ' ') return 0;
  for (i = n; i < lb; i++) if (b[i] != ' ') return 0;
  return 1;
}
static void f77_strcpy(char *d, long ld, const char *s, long ls) {
  long n = ld < ls ? ld : ls; memmove(d, s, (size_t)n); if (ld > n) memset(d + n, ' ', (size_t)(ld - n));
}
static int f77_len_trim(const char *a, long la) { while (la > 0 && a[la - 1] == ' ') la--; return (int)la; }
static int f77_index(const char *a, long la, const char *b, long lb) {
  for (long i = 0; i + lb <= la; i++) if (!memcmp(a + i, b, (size_t)lb)) return (int)i + 1; return 0;
}
static void f77_unsupported(const char *what) { fprintf(stderr, "f77c: unsupported construct executed: %s\n", what); abort(); }
static void f77_stop(void) { fprintf(stderr, "f77c: STOP\n"); exit(1); }
extern int f77_trace_on;                                       /* oracle/ref_stubs.c */
extern void f77_trace(const char *unit, int n, ...);
"""


class Translator:
    def __init__(self, root, include_dirs, include_map=None, defines=()):
        self.root = root
        self.reader = Reader(include_dirs, include_map, defines)
        self.units = {}           # name -> info
        self.common_size = {}
        self.common_views = {}
        self.common_used = set()
        self.externs = {}         # name -> return C type as seen by callers
        self.defined = {}
        self.known_units = set()
        self.failed = {}
        self.trace_units = set()  # routines whose write(6,...) statements report their numeric items to f77_trace

    def add_file(self, path, only=None, skip=()):
        for u in split_units(self.reader.read(path)):
            if u["name"] in skip or (only is not None and u["name"] not in only):
                continue
            if u["kind"] in ("program", "blockdata"):
                continue
            if u["name"] not in self.units:      # first definition wins
                self.units[u["name"]] = u
                self.known_units.add(u["name"])

    def unit_kind(self, n):
        return self.units[n]["kind"] if n in self.units else None

    def note_extern(self, n, rt, unit):
        if rt is None:
            rt = CTYPE[unit.typeof_sym(unit.sym(n))]
        self.externs.setdefault(n, rt)

    def translate(self, roots, stop_at=()):
        """Translates `roots` and everything they call that is known; returns (C text, missing callees)."""
        done, order, missing = {}, [], set()
        work = list(roots)
        while work:
            n = work.pop()
            if n in done:
                continue
            if n not in self.units:
                missing.add(n)
                continue
            if n in stop_at:
                missing.add(n)
                continue
            try:
                u = Unit(self, self.units[n])
                code = u.translate()
            except (TranslationError, SyntaxError) as e:
                # not translatable: a stub that aborts loudly if it is ever executed
                self.failed[n] = str(e)
                rt = "double" if self.units[n]["kind"] == "function" else "void"
                self.defined[n] = rt
                msg = (n + ": " + str(e).splitlines()[0]).replace("\\", "/").replace('"', "'")
                done[n] = [f'{rt} {n}_() {{ f77_unsupported("{msg}"); {"return 0;" if rt != "void" else ""} }}', ""]
                order.append(n)
                continue
            done[n] = code
            order.append(n)
            for c in sorted(u.calls):
                if c not in done:
                    work.append(c)
        out = ["/* generated by oracle/f77c.py from the reference sources; do not commit */", RUNTIME]
        for blk in sorted(self.common_size):
            out.append(f"char cb_{blk}[{max(self.common_size[blk], 1)}] __attribute__((aligned(64)));")
        out.append("")
        for n, rt in sorted(self.externs.items()):
            rt = self.defined.get(n, rt)
            out.append(f"{rt} {n}_();")
        out.append("")
        for n in order:
            out.extend(done[n])
        return "\n".join(out), sorted(missing - set(done))

    def common_map(self):
        """{var: {block, offset, type, elsize, dims, lows}}: for a name that different routines place differently (local
        COMMON declarations with their own names), the placement used by the most routines wins (the include files')."""
        votes = {}
        for blk, views in self.common_views.items():
            for unit, lst in views.items():
                for n, off, t, es, dims, lows in lst:
                    key = (blk, off, t, es, tuple(dims), tuple(lows))
                    votes.setdefault(n, {}).setdefault(key, 0)
                    votes[n][key] += 1
        m = {}
        for n, d in votes.items():
            (blk, off, t, es, dims, lows), _ = max(d.items(), key=lambda kv: kv[1])
            m[n] = dict(block=blk, offset=off, type=t, elsize=es, dims=list(dims), lows=list(lows))
        return m

    def common_alt(self):
        """{unit: {var: entry}} for the placements that differ from common_map()'s winner (routine-local COMMON views)."""
        m = self.common_map()
        alt = {}
        for blk, views in self.common_views.items():
            for unit, lst in views.items():
                for n, off, t, es, dims, lows in lst:
                    e = dict(block=blk, offset=off, type=t, elsize=es, dims=list(dims), lows=list(lows))
                    if m[n] != e:
                        alt.setdefault(unit, {})[n] = e
        return alt


if __name__ == "__main__":
    print(__doc__)
