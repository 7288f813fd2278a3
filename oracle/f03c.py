#!/usr/bin/env python
"""f03c -- a translator from the Fortran 2003 subset used by @mrg37-080A.f03 to C99.

TEST INFRASTRUCTURE.  The image has no Fortran compiler (gfortran, flang, nvfortran, f2c: none), so the
reference cannot be built the usual way.  This module is a small *language implementation*, not a
restatement of any algorithm: it knows nothing about fulmov or particles.  oracle/build_ref.py runs it over
the reference's own source file where it lies (/root/reference/@mrg37-080A.f03) and compiles the generated C
with gcc into oracle/_ref/ (git-ignored, never committed: it is derived from GPL-3.0 reference source).  The
resulting library IS the reference's code for the routines it contains, executed statement by statement with
the reference's own expression order, and it is what pins oracle/fulmov_oracle.c (tests/test_ref_pin.py).

Subset: free-form source, `&` continuations, `!` comments; program units subroutine / function;
declarations real(C_DOUBLE) | real(C_float) | integer(C_INT) | logical | character (ignored) with
dimension(...) / save / parameter attributes and initialisers, COMMON, PARAMETER, DATA (scalars),
include files; executable statements: assignment (scalar, whole-array fill), block and logical IF, DO
(with or without label, with step), DO WHILE, labelled CONTINUE, GO TO, CALL, RETURN, CYCLE, EXIT, STOP;
EQUIVALENCE of local arrays overlaid from their first elements (the one form on the path, F:6107); ENTRY without
arguments (`entry prefld` inside emfild, F:3820: one body function with a selector, one wrapper per callable
name); mpi_allreduce / mpi_allgather / mpi_isend / mpi_irecv / mpi_wait map onto the simulated ranks of
oracle/ref_runtime.c; unformatted sequential I/O on a numbered unit (`write(12) list`, `read(12) list` with implied
DO lists, `open(...,form='unformatted')`, `close`: the restart file of restrt) becomes record calls of that run time.
Formatted I/O statements (write/read/open/close/print/format/rewind on the log and plot units) are dropped.  Expressions follow Fortran
typing: integer division truncates, default-real literals (no `d` exponent) are single precision,
mixed-mode promotion as in Fortran (which C's usual arithmetic conversions reproduce), x**n by the
multiplication chain gcc/gfortran use (__powidf2 order), left-to-right association of equal-precedence
operators, parentheses kept.  All arguments are passed by reference, as gfortran does.

Semantics that matter for bit-exactness and how they are kept:
  * no FMA contraction, no re-association: the C is compiled with -O2 -ffp-contract=off -fwrapv and without
    -ffast-math, for baseline x86-64 (the documented build line `mpif90 -mcmodel=medium -O2`, F:100, has
    no -march either, so gfortran emits no FMA);
  * int32 wrap-around in the LCG (iand(lambda*ir, mask), F:9301): -fwrapv;
  * compile-time sizes (param_080A.h: mx,my,mz,np0,npc and what derives from them) become run-time
    globals so that one library serves every test size -- editing param_080A.h per run size is how the
    reference is used (P:12-17);
  * COMMON blocks are storage sequences: every unit's view is laid out from its own declaration, so
    units that name the members differently still alias the same bytes;
  * one simulated MPI rank = one thread with private COMMON storage (oracle/ref_runtime.c).
"""
import re
import sys

# ----------------------------------------------------------------------------------------------
# source reading
# ----------------------------------------------------------------------------------------------


def strip_comment(line):
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


def lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


def read_statements(path, include_dirs=()):
    """-> list of (lineno, label or None, text) logical statements, lower-cased outside strings."""
    stmts = []
    cur, cur_line = "", 0
    with open(path, errors="replace") as f:
        for n, raw in enumerate(f, 1):
            s = strip_comment(raw.rstrip("\n")).replace("\t", " ")
            if not s.strip():
                continue
            t = s.strip()
            if cur:
                if t.startswith("&"):
                    t = t[1:]
                cur += " " + t
            else:
                cur, cur_line = t, n
            if cur.endswith("&"):
                cur = cur[:-1].rstrip()
                continue
            stmts.append((cur_line, cur))
            cur = ""
    out = []
    for n, s in stmts:
        s = lower_outside_strings(s)
        m = re.match(r"^(\d+)\s+(.*)$", s)
        label = None
        if m:
            label, s = m.group(1), m.group(2)
        out.append((n, label, s.strip()))
    return out


# ----------------------------------------------------------------------------------------------
# lexer / expression parser
# ----------------------------------------------------------------------------------------------
DOTOPS = {".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".and.": "&&", ".or.": "||",
          ".not.": "!", ".true.": "T", ".false.": "F", ".eqv.": "eqv", ".neqv.": "neqv"}
TOK = re.compile(r"""
   (?P<dot>\.(?:eq|ne|lt|le|gt|ge|and|or|not|true|false|eqv|neqv)\.)
 | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?)
 | (?P<name>[a-z_][a-z0-9_]*)
 | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
 | (?P<op>\*\*|==|/=|<=|>=|//|::|=>|[-+*/(),=<>:%])
 | (?P<ws>\s+)
""", re.X)


class Tok:
    __slots__ = ("k", "v")

    def __init__(self, k, v):
        self.k, self.v = k, v

    def __repr__(self):
        return "%s:%s" % (self.k, self.v)


def lex(s):
    toks, i = [], 0
    while i < len(s):
        m = TOK.match(s, i)
        if not m:
            raise SyntaxError("cannot lex %r at %r" % (s, s[i:i + 20]))
        k = m.lastgroup
        v = m.group(k)
        if k == "num":
            # "1.eq.2": the dot belongs to the operator, not to the number
            if "." in v and not re.search(r"[ed]", v) and v.endswith("."):
                m2 = re.match(r"\.(?:eq|ne|lt|le|gt|ge|and|or|not|eqv|neqv)\.", s[m.end() - 1:])
                if m2:
                    v = v[:-1]
                    toks.append(Tok("num", v))
                    i = m.end() - 1
                    continue
            # "1.d0" lexes whole; "3.e" cannot occur
        if k != "ws":
            toks.append(Tok(k, v))
        i = m.end()
    return toks


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else Tok("eof", "")

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def accept(self, v):
        if self.peek().v == v and self.peek().k in ("op", "dot", "name"):
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise SyntaxError("expected %r, got %r in %r" % (v, self.peek(), self.t))

    def at_end(self):
        return self.i >= len(self.t)

    # precedence climbing
    def expr(self):
        return self.p_eqv()

    def p_eqv(self):
        e = self.p_or()
        while self.peek().k == "dot" and self.peek().v in (".eqv.", ".neqv."):
            op = self.next().v
            e = ("bin", "==" if op == ".eqv." else "!=", e, self.p_or())
        return e

    def p_or(self):
        e = self.p_and()
        while self.peek().k == "dot" and self.peek().v == ".or.":
            self.next()
            e = ("bin", "||", e, self.p_and())
        return e

    def p_and(self):
        e = self.p_not()
        while self.peek().k == "dot" and self.peek().v == ".and.":
            self.next()
            e = ("bin", "&&", e, self.p_not())
        return e

    def p_not(self):
        if self.peek().k == "dot" and self.peek().v == ".not.":
            self.next()
            return ("un", "!", self.p_not())
        return self.p_rel()

    def p_rel(self):
        e = self.p_add()
        t = self.peek()
        rel = None
        if t.k == "dot" and t.v in (".eq.", ".ne.", ".lt.", ".le.", ".gt.", ".ge."):
            rel = DOTOPS[t.v]
        elif t.k == "op" and t.v in ("==", "/=", "<", "<=", ">", ">="):
            rel = "!=" if t.v == "/=" else t.v
        if rel:
            self.next()
            e = ("bin", rel, e, self.p_add())
        return e

    def p_add(self):
        t = self.peek()
        if t.k == "op" and t.v in "+-" and len(t.v) == 1:
            self.next()
            e = self.p_mul()
            if t.v == "-":
                e = ("un", "-", e)
        else:
            e = self.p_mul()
        while self.peek().k == "op" and self.peek().v in ("+", "-"):
            op = self.next().v
            e = ("bin", op, e, self.p_mul())
        return e

    def p_mul(self):
        e = self.p_pow()
        while self.peek().k == "op" and self.peek().v in ("*", "/"):
            op = self.next().v
            e = ("bin", op, e, self.p_pow())
        return e

    def p_pow(self):
        e = self.p_primary()
        if self.peek().k == "op" and self.peek().v == "**":
            self.next()
            # right associative; the exponent may carry a sign
            t = self.peek()
            if t.k == "op" and t.v in ("+", "-"):
                self.next()
                r = self.p_pow()
                if t.v == "-":
                    r = ("un", "-", r)
            else:
                r = self.p_pow()
            e = ("bin", "**", e, r)
        return e

    def p_primary(self):
        t = self.next()
        if t.k == "num":
            return ("num", t.v)
        if t.k == "str":
            return ("str", t.v)
        if t.k == "dot" and t.v in (".true.", ".false."):
            return ("log", t.v == ".true.")
        if t.k == "op" and t.v == "(":
            e = self.expr()
            self.expect(")")
            return ("par", e)
        if t.k == "name":
            if self.peek().k == "op" and self.peek().v == "(":
                self.next()
                args = []
                if not (self.peek().k == "op" and self.peek().v == ")"):
                    while True:
                        args.append(self.arg())
                        if not self.accept(","):
                            break
                self.expect(")")
                return ("call", t.v, args)
            return ("var", t.v)
        if t.k == "op" and t.v in ("+", "-"):
            e = self.p_primary()
            return ("un", "-", e) if t.v == "-" else e
        raise SyntaxError("unexpected token %r in %r" % (t, self.t))

    def arg(self):
        # array section a:b or ":" (only whole-dimension sections are supported, and only in declarations)
        if self.peek().k == "op" and self.peek().v == ":":
            self.next()
            return ("colon",)
        e = self.expr()
        if self.peek().k == "op" and self.peek().v == ":":
            self.next()
            hi = self.expr()
            return ("range", e, hi)
        return e


def parse_expr(s):
    p = Parser(lex(s))
    e = p.expr()
    if not p.at_end():
        raise SyntaxError("trailing tokens in expression %r" % s)
    return e


def split_top(s, sep=","):
    """split at separators outside parentheses and strings"""
    out, depth, q, cur = [], 0, None, []
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur).strip())
    return out


def match_paren(s, i):
    """index of the ')' matching the '(' at s[i]"""
    depth, q = 0, None
    for j in range(i, len(s)):
        ch = s[j]
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return j
    raise SyntaxError("unbalanced parentheses in %r" % s)


# ----------------------------------------------------------------------------------------------
# program units and symbols
# ----------------------------------------------------------------------------------------------
CTYPE = {"int": "int", "double": "double", "real": "float", "logical": "int"}
BYTES = {"int": 4, "double": 8, "real": 4, "logical": 4}
RANK = {"logical": 0, "int": 1, "real": 2, "double": 3}


class Sym:
    def __init__(self, name):
        self.name = name
        self.type = None
        self.dims = None          # list of (lo_ast, hi_ast) or None
        self.kind = "local"       # local | dummy | common | param | result
        self.block = None
        self.value = None         # parameter value AST / initialiser AST
        self.save = False
        self.is_char = False

    def __repr__(self):
        return "Sym(%s,%s,%s,%s)" % (self.name, self.type, self.kind, "arr" if self.dims else "scalar")


class Unit:
    def __init__(self, kind, name, args, prefix_type, line):
        self.kind, self.name, self.args, self.line = kind, name, args, line
        self.prefix_type = prefix_type
        self.stmts = []
        self.sym = {}
        self.order = []           # declaration order of symbols
        self.commons = []         # list of (block, [names]) in declaration order
        self.equiv = {}           # alias -> owner (EQUIVALENCE of local arrays from their first elements)
        self.entries = []         # (name, args) of ENTRY statements


UNIT_RE = re.compile(r"^(?:(integer|real|double precision|logical)\s*(?:\(([^)]*)\)|\*\s*(\d+))?\s+)?(subroutine|function)\s+([a-z_0-9]+)\s*(?:\((.*)\))?\s*$")


def split_units(stmts):
    units, cur = [], None
    for n, label, s in stmts:
        if cur is None:
            m = UNIT_RE.match(s)
            if m:
                ptype = None
                if m.group(1):
                    ptype = type_from_spec(m.group(1), m.group(2) or (("*" + m.group(3)) if m.group(3) else ""))
                args = [a.strip() for a in (m.group(6) or "").split(",") if a.strip()]
                cur = Unit(m.group(4), m.group(5), args, ptype, n)
            elif re.match(r"^(program|block\s*data|module)\b", s):
                cur = Unit("other", s, [], None, n)
            continue
        if re.match(r"^end\s*(subroutine|function|program|block\s*data|module)?(\s+[a-z_0-9]+)?\s*$", s) and not re.match(r"^end\s*(do|if)\b", s):
            units.append(cur)
            cur = None
            continue
        cur.stmts.append((n, label, s))
    return units


def type_from_spec(base, kind):
    base = base.strip()
    kind = (kind or "").replace(" ", "").lower()
    if base == "integer":
        return "int"
    if base == "logical":
        return "logical"
    if base == "double precision":
        return "double"
    if base == "real":
        if kind in ("c_double", "8", "*8", "kind=8", "kind=c_double"):
            return "double"
        if kind in ("", "c_float", "4", "*4", "kind=4", "kind=c_float"):
            return "real"
        raise SyntaxError("unknown real kind %r" % kind)
    raise SyntaxError("unknown type %r" % base)


DECL_RE = re.compile(r"^(integer|real|double precision|logical|character|complex)\b")
IO_RE = re.compile(r"^(write|read|open|close|print|format|rewind|backspace|flush)\b\s*[(\*'\"]?")


class Translator:
    def __init__(self, path, include_dirs, runtime_params, want, stubs=()):
        self.stubs = set(stubs)
        self.unf_units = set()     # unit numbers opened with form='unformatted' somewhere in the translated units
        self.path = path
        self.include_dirs = include_dirs
        self.runtime_params = set(runtime_params)     # parameter names that become run-time globals
        self.units = {}
        self.include_cache = {}
        stmts = read_statements(path)
        for u in split_units(stmts):
            if u.kind in ("subroutine", "function"):
                self.units[u.name] = u
        # ENTRY without arguments (`entry prefld` inside emfild, F:3820): a callable name that jumps into its host unit
        self.entry_host = {}
        for u in list(self.units.values()):
            if u.kind != "subroutine":
                continue
            for n, label, t in u.stmts:
                m = re.match(r"^entry\s+([a-z_][a-z0-9_]*)\s*(\(\s*\))?$", t)
                if m and m.group(1) not in self.units:
                    e = Unit("subroutine", m.group(1), [], None, n)
                    self.units[e.name] = e
                    self.entry_host[e.name] = u.name
        self.want = list(want)
        self.param_globals = {}      # name -> (type, value AST) of every PARAMETER seen in include files (emitted once)
        self.param_order = []
        self.blocks = {}             # block -> first declaring unit's member list [(name,type,dims)]
        self.block_order = []
        self.out = []
        self.externals = set()

    # -- declarations -----------------------------------------------------------------------------
    def include_stmts(self, fname):
        if fname not in self.include_cache:
            import os
            for d in self.include_dirs:
                p = os.path.join(d, fname)
                if os.path.exists(p):
                    self.include_cache[fname] = read_statements(p)
                    break
            else:
                self.include_cache[fname] = None
        return self.include_cache[fname]

    def sym(self, u, name):
        if name not in u.sym:
            u.sym[name] = Sym(name)
            u.order.append(name)
        return u.sym[name]

    def parse_dims(self, s):
        dims = []
        for d in split_top(s):
            if ":" in d:
                parts = split_top(d, ":")
                lo, hi = parts[0], parts[1]
                dims.append((parse_expr(lo), parse_expr(hi) if hi.strip() != "*" else None))
            else:
                dims.append((("num", "1"), parse_expr(d) if d.strip() != "*" else None))
        return dims

    def parse_entity(self, u, ent, typ, attr_dims, is_param, is_save, from_include):
        m = re.match(r"^([a-z_][a-z0-9_]*)\s*(.*)$", ent)
        if not m:
            raise SyntaxError("bad entity %r" % ent)
        name, rest = m.group(1), m.group(2).strip()
        dims, init = attr_dims, None
        if rest.startswith("("):
            j = match_paren(rest, 0)
            dims = self.parse_dims(rest[1:j])
            rest = rest[j + 1:].strip()
        if rest.startswith("*"):           # character length: name*29
            rest = re.sub(r"^\*\s*\d+", "", rest).strip()
        if rest.startswith("="):
            init = parse_expr(rest[1:].strip())
        s = self.sym(u, name)
        if typ == "char":
            s.is_char = True
        elif typ:
            s.type = typ
        if dims:
            s.dims = dims
        if is_save or init is not None:
            s.save = True
        if init is not None:
            s.value = init
        if is_param:
            s.kind = "param"
        return s

    def parse_decl(self, u, s, from_include=False):
        """returns True when the statement was a declaration"""
        if s.startswith("use ") or s.startswith("use,") or s.startswith("implicit "):
            return True
        m = re.match(r"^include\s+['\"]([^'\"]+)['\"]", s)
        if m:
            inc = self.include_stmts(m.group(1))
            if inc is None:
                if m.group(1) != "mpif.h":
                    raise SyntaxError("include file %r not found" % m.group(1))
                return True
            for n, label, t in inc:
                if not self.parse_decl(u, t, True):
                    raise SyntaxError("executable statement in include file: %r" % t)
            return True
        if s.startswith("common"):
            rest = s[6:].strip()
            m = re.match(r"^/\s*([a-z_0-9]*)\s*/\s*(.*)$", rest)
            if not m:
                raise SyntaxError("bad COMMON %r" % s)
            blk, lst = m.group(1) or "blank", m.group(2)
            names = []
            for ent in split_top(lst):
                sy = self.parse_entity(u, ent, None, None, False, False, from_include)
                sy.kind, sy.block = "common", blk
                names.append(sy.name)
            u.commons.append((blk, names))
            return True
        if s.startswith("parameter"):
            rest = s[9:].strip()
            assert rest.startswith("(") and rest.endswith(")"), s
            for ent in split_top(rest[1:-1]):
                name, _, val = ent.partition("=")
                name = name.strip()
                sy = self.sym(u, name)
                sy.kind = "param"
                if sy.is_char or val.strip().startswith(("'", '"')):
                    sy.is_char = True
                    continue
                sy.value = parse_expr(val.strip())
            return True
        if s.startswith("dimension"):
            for ent in split_top(s[9:].strip()):
                self.parse_entity(u, ent, None, None, False, False, from_include)
            return True
        if re.match(r"^(external|intrinsic)\b", s):
            return True
        if s.startswith("save"):
            for nme in split_top(s[4:].replace("::", "").strip()):
                if nme and not nme.startswith("/"):
                    self.sym(u, nme).save = True
            return True
        if s.startswith("equivalence"):
            # the one form the reference uses on the path: local arrays of one type overlaid from their first elements,
            # equivalence (w0(1),w1(1,1)) -- the later names become aliases of the first
            rest = s[len("equivalence"):].strip()
            while rest:
                if not rest.startswith("("):
                    raise SyntaxError("EQUIVALENCE form is not supported: %r" % s)
                j = match_paren(rest, 0)
                ents = split_top(rest[1:j])
                rest = rest[j + 1:].lstrip().lstrip(",").lstrip()
                names = []
                for ent in ents:
                    ent = ent.strip()
                    mm = re.match(r"^([a-z_][a-z0-9_]*)\s*(\((.*)\))?$", ent)
                    if not mm or (mm.group(3) and any(x.strip() != "1" for x in mm.group(3).split(","))):
                        raise SyntaxError("EQUIVALENCE with an offset is not supported: %r" % s)
                    names.append(mm.group(1))
                for nme in names[1:]:
                    u.equiv[nme] = names[0]
            return True
        if s.startswith("data "):
            # data name/value/ [, name/value/ ...]  (scalars only)
            for mm in re.finditer(r"([a-z_][a-z0-9_]*)\s*/\s*([^/]+)/", s[5:]):
                sy = self.sym(u, mm.group(1))
                sy.value = parse_expr(mm.group(2).strip())
                sy.save = True
            return True
        m = DECL_RE.match(s)
        if m and not re.match(r"^(real|integer|logical)\s*=", s):
            base = m.group(1)
            rest = s[m.end():].lstrip()
            kind = ""
            if rest.startswith("("):
                j = match_paren(rest, 0)
                kind, rest = rest[1:j], rest[j + 1:].lstrip()
            elif rest.startswith("*"):
                mm = re.match(r"^\*\s*(\d+)", rest)
                kind, rest = "*" + mm.group(1), rest[mm.end():].lstrip()
            typ = "char" if base == "character" else type_from_spec(base, kind)
            attr_dims, is_param, is_save = None, False, False
            if "::" in rest:
                attrs, _, rest = rest.partition("::")
                for a in split_top(attrs.strip().lstrip(",")):
                    a = a.strip()
                    if a.startswith("dimension"):
                        i0 = a.index("(")
                        attr_dims = self.parse_dims(a[i0 + 1:match_paren(a, i0)])
                    elif a == "parameter":
                        is_param = True
                    elif a == "save":
                        is_save = True
                    elif a.startswith("intent") or a in ("", "target", "optional", "value"):
                        pass
                    else:
                        raise SyntaxError("unsupported attribute %r in %r" % (a, s))
            elif rest.startswith(","):
                # "real(C_DOUBLE),dimension(np0) x" never occurs without "::"; be strict
                raise SyntaxError("attribute list without '::' in %r" % s)
            for ent in split_top(rest.strip()):
                self.parse_entity(u, ent, typ, attr_dims, is_param, is_save, from_include)
            return True
        return False

    # -- typing -----------------------------------------------------------------------------------
    def num_type(self, v):
        if re.search(r"d", v):
            return "double"
        if "." in v or "e" in v:
            return "real"
        return "int"

    def etype(self, u, e):
        k = e[0]
        if k == "num":
            return self.num_type(e[1])
        if k == "log":
            return "logical"
        if k == "str":
            return "char"
        if k == "par":
            return self.etype(u, e[1])
        if k == "un":
            return "logical" if e[1] == "!" else self.etype(u, e[2])
        if k == "bin":
            op = e[1]
            if op in ("==", "!=", "<", "<=", ">", ">=", "&&", "||"):
                return "logical"
            a, b = self.etype(u, e[2]), self.etype(u, e[3])
            if op == "**" and RANK[b] <= 1:
                return a
            return a if RANK[a] >= RANK[b] else b
        if k == "var":
            return self.var_type(u, e[1])
        if k == "call":
            name = e[1]
            s = u.sym.get(name)
            if s and s.dims:
                return s.type or implicit_type(name)
            return self.func_type(u, name, e[2])
        raise SyntaxError("etype %r" % (e,))

    def var_type(self, u, name):
        s = u.sym.get(name)
        if s and s.type:
            return s.type
        if name.startswith("mpi_"):
            return "int"
        if name in self.param_globals:
            return self.param_globals[name][0]
        return implicit_type(name)

    INTRINSIC_TYPES = {"int": "int", "ifix": "int", "idint": "int", "nint": "int", "idnint": "int", "iand": "int", "ior": "int",
                       "ieor": "int", "ishft": "int", "iabs": "int", "dble": "double", "dfloat": "double", "float": "real",
                       "sngl": "real", "dabs": "double", "dsqrt": "double", "dexp": "double", "dlog": "double", "dsin": "double",
                       "dcos": "double", "datan": "double", "datan2": "double", "dtanh": "double", "dmax1": "double",
                       "dmin1": "double", "amax1": "real", "amin1": "real", "max0": "int", "min0": "int", "dsign": "double",
                       "isign": "int", "dmod": "double", "amod": "real", "alog": "real", "alog10": "real", "dlog10": "double"}
    GENERIC = {"abs", "sqrt", "exp", "log", "log10", "sin", "cos", "tan", "atan", "atan2", "tanh", "cosh", "sinh", "asin",
               "acos", "mod", "min", "max", "sign", "aint", "anint"}

    def func_type(self, u, name, args):
        if name in self.INTRINSIC_TYPES:
            return self.INTRINSIC_TYPES[name]
        if name == "real":
            return "real" if len(args) == 1 else "double"
        if name in self.GENERIC:
            ts = [self.etype(u, a) for a in args]
            t = ts[0]
            for x in ts[1:]:
                t = t if RANK[t] >= RANK[x] else x
            return t
        # external function: declared type in the caller, or the callee's result type
        s = u.sym.get(name)
        if s and s.type:
            return s.type
        if name in self.units:
            cu = self.units[name]
            return cu.prefix_type or implicit_type(name)
        return implicit_type(name)

    # -- expression emission ----------------------------------------------------------------------
    def cnum(self, v):
        t = self.num_type(v)
        if t == "int":
            return v
        if t == "double":
            x = v.replace("d", "e")
            if "." not in x.split("e")[0]:
                x = x.replace("e", ".0e", 1)
            return x
        x = v
        if "." not in x.split("e")[0]:
            x = x.replace("e", ".0e", 1) if "e" in x else x + ".0"
        if x.endswith("."):
            x += "0"
        x = re.sub(r"\.e", ".0e", x)
        if x.startswith("."):
            x = "0" + x
        return x + "f"

    def cvar(self, name):
        return name + "_"

    def emit_expr(self, u, e, ctx):
        k = e[0]
        if k == "num":
            return self.cnum(e[1])
        if k == "log":
            return "1" if e[1] else "0"
        if k == "par":
            return "(" + self.emit_expr(u, e[1], ctx) + ")"
        if k == "un":
            return "(" + e[1] + self.emit_expr(u, e[2], ctx) + ")"
        if k == "bin":
            op = e[1]
            if op == "**":
                return self.emit_pow(u, e[2], e[3], ctx)
            a, b = self.emit_expr(u, e[2], ctx), self.emit_expr(u, e[3], ctx)
            if op in ("&&", "||"):
                return "(%s %s %s)" % (a, op, b)
            return "(%s %s %s)" % (a, op, b)
        if k == "var":
            return self.emit_var(u, e[1], ctx)
        if k == "call":
            name = e[1]
            s = u.sym.get(name)
            if s and s.dims:
                return "%s[%s]" % (self.cvar(name), self.emit_index(u, s, e[2], ctx))
            return self.emit_func(u, name, e[2], ctx)
        raise SyntaxError("emit %r" % (e,))

    def emit_var(self, u, name, ctx):
        s = u.sym.get(name)
        if s is None:
            if name.startswith("mpi_"):
                return "REF_" + name.upper()
            if name in self.param_globals:
                return self.cvar(name)
            raise SyntaxError("undeclared variable %r in %s" % (name, u.name))
        if s.kind == "param":
            return self.cvar(name)
        if s.dims:
            return self.cvar(name)          # whole array (only as an actual argument)
        if s.kind in ("dummy", "common"):
            return "(*%s)" % self.cvar(name)
        return self.cvar(name)

    def emit_index(self, u, s, args, ctx):
        if len(args) != len(s.dims):
            raise SyntaxError("rank mismatch for %s in %s" % (s.name, u.name))
        n = self.cvar(s.name)
        # offset = sum_k (i_k - lo_k) * stride_k ; strides and the constant part are locals computed at entry
        parts = []
        for k, a in enumerate(args):
            if a[0] in ("colon", "range"):
                raise SyntaxError("array sections are not supported (%s in %s)" % (s.name, u.name))
            ix = self.emit_expr(u, a, ctx)
            parts.append(ix if k == 0 else "%s_S%d*(long)%s" % (n, k, ix))
        return "%s_O + %s" % (n, " + ".join(parts))

    def emit_pow(self, u, base, ex, ctx):
        tb, te = self.etype(u, base), self.etype(u, ex)
        b = self.emit_expr(u, base, ctx)
        if RANK[te] <= 1:
            # small literal exponents expand to the multiplication chain gcc uses; others go through the powi helpers
            if tb == "int":
                return "ref_ipow(%s, %s)" % (b, self.emit_expr(u, ex, ctx))
            if tb == "real":
                return "ref_powif(%s, %s)" % (b, self.emit_expr(u, ex, ctx))
            return "ref_powi(%s, %s)" % (b, self.emit_expr(u, ex, ctx))
        if tb == "real" and te == "real":
            return "powf(%s, %s)" % (b, self.emit_expr(u, ex, ctx))
        return "pow((double)%s, (double)%s)" % (b, self.emit_expr(u, ex, ctx))

    MATH1 = {"sqrt": "sqrt", "exp": "exp", "log": "log", "log10": "log10", "sin": "sin", "cos": "cos", "tan": "tan",
             "atan": "atan", "tanh": "tanh", "cosh": "cosh", "sinh": "sinh", "asin": "asin", "acos": "acos",
             "dsqrt": "sqrt", "dexp": "exp", "dlog": "log", "dsin": "sin", "dcos": "cos", "datan": "atan", "dtanh": "tanh",
             "alog": "log", "alog10": "log10", "dlog10": "log10"}

    def emit_func(self, u, name, args, ctx):
        ea = [self.emit_expr(u, a, ctx) for a in args]
        ts = [self.etype(u, a) for a in args]
        if name in ("int", "ifix", "idint"):
            return "((int)(%s))" % ea[0]
        if name in ("nint", "idnint"):
            return "((int)lround(%s))" % ea[0] if ts[0] == "double" else "((int)lroundf(%s))" % ea[0]
        if name in ("dble", "dfloat"):
            return "((double)(%s))" % ea[0]
        if name in ("float", "sngl") or (name == "real" and len(args) == 1):
            return "((float)(%s))" % ea[0]
        if name in ("abs", "dabs", "iabs"):
            t = ts[0]
            return ("abs(%s)" if t == "int" else ("fabsf(%s)" if t == "real" else "fabs(%s)")) % ea[0]
        if name in self.MATH1:
            f = self.MATH1[name]
            if ts[0] == "real" and name in self.GENERIC:
                return "%sf(%s)" % (f, ea[0])
            return "%s((double)(%s))" % (f, ea[0])
        if name in ("atan2", "datan2"):
            return "atan2((double)(%s), (double)(%s))" % (ea[0], ea[1])
        if name in ("mod", "dmod", "amod"):
            t = self.func_type(u, "mod", args)
            if t == "int":
                return "((%s) %% (%s))" % (ea[0], ea[1])
            return ("fmodf(%s, %s)" if t == "real" else "fmod(%s, %s)") % (ea[0], ea[1])
        if name == "iand":
            return "((%s) & (%s))" % (ea[0], ea[1])
        if name == "ior":
            return "((%s) | (%s))" % (ea[0], ea[1])
        if name == "ieor":
            return "((%s) ^ (%s))" % (ea[0], ea[1])
        if name in ("min", "max", "dmin1", "dmax1", "amin1", "amax1", "min0", "max0"):
            t = self.func_type(u, name if name in self.INTRINSIC_TYPES else "min", args)
            base = "min" if "min" in name else "max"
            fn = {"int": "ref_i%s", "real": "ref_f%s", "double": "ref_d%s"}[t] % base
            out = ea[0]
            for x in ea[1:]:
                out = "%s(%s, %s)" % (fn, out, x)
            return out
        if name in ("sign", "dsign", "isign"):
            t = self.func_type(u, "sign", args)
            return {"int": "ref_isign", "real": "ref_fsign", "double": "ref_dsign"}[t] + "(%s, %s)" % (ea[0], ea[1])
        if name in ("aint",):
            return "trunc(%s)" % ea[0]
        if name in ("anint",):
            return "round(%s)" % ea[0]
        # external function of the translated set
        if name not in self.units:
            raise SyntaxError("unknown function %r called in %s" % (name, u.name))
        self.externals.add(name)
        self.ensure_decls(self.units[name])
        return "%s_f(%s)" % (name, ", ".join(self.emit_actual(u, a, self.units[name], i, ctx) for i, a in enumerate(args)))

    def emit_actual(self, u, a, callee, i, ctx):
        """an actual argument, by reference"""
        while a[0] == "par" and a[1][0] in ("var",):
            a = a[1]
        if a[0] == "var":
            s = u.sym.get(a[1])
            if s is not None and s.kind != "param":
                if s.dims:
                    return "(void*)%s" % self.cvar(a[1])
                if s.kind in ("dummy", "common"):
                    return "(void*)%s" % self.cvar(a[1])
                return "(void*)&%s" % self.cvar(a[1])
        if a[0] == "call":
            s = u.sym.get(a[1])
            if s and s.dims:
                return "(void*)&%s[%s]" % (self.cvar(a[1]), self.emit_index(u, s, a[2], ctx))
        if a[0] == "str":
            return "(void*)0"
        t = self.etype(u, a)
        # the temporary takes the DUMMY's type when the callee is known (gfortran would pass the bits as they are;
        # the sources on the path always agree, and a mismatch is reported)
        if callee is not None and i < len(callee.args):
            ds = callee.sym.get(callee.args[i])
            dt = (ds.type if ds and ds.type else implicit_type(callee.args[i]))
            if dt != t and not (ds and ds.is_char):
                raise SyntaxError("argument %d of %s: actual is %s, dummy is %s (in %s)" % (i + 1, callee.name, t, dt, u.name))
        return "(void*)(%s[]){%s}" % (CTYPE[t], self.emit_expr(u, a, ctx))

    # -- statements -------------------------------------------------------------------------------
    def translate_unit(self, u):
        # 1. declarations
        body_start = 0
        for idx, (n, label, s) in enumerate(u.stmts):
            try:
                if label is None and self.parse_decl(u, s):
                    body_start = idx + 1
                    continue
            except SyntaxError as ex:
                raise SyntaxError("%s:%d: %s" % (u.name, n, ex))
            # format statements may sit among declarations
            if IO_RE.match(s) and s.startswith("format"):
                body_start = idx + 1
                continue
            break
        for a in u.args:
            s = self.sym(u, a)
            s.kind = "dummy"
        if u.kind == "function":
            s = self.sym(u, u.name)
            s.kind = "result"
            if u.prefix_type:
                s.type = u.prefix_type
        for name, s in u.sym.items():
            if s.type is None and not s.is_char:
                s.type = implicit_type(name)
        # parameters from include files become globals (emitted once); local parameters stay local constants
        # 2. code
        L = []
        ctx = {"labels": set(), "do_stack": [], "tmp": 0}
        rett = CTYPE[u.sym[u.name].type] if u.kind == "function" else "void"
        params = []
        for a in u.args:
            s = u.sym[a]
            if s.is_char:
                params.append("void *%s" % self.cvar(a))
            else:
                params.append("%s *%s" % (CTYPE[s.type], self.cvar(a)))
        hosted = [e for e, h in self.entry_host.items() if h == u.name]
        if hosted:
            # wrappers first (the first line of the text is the unit's prototype): every callable name enters one body
            # function with a selector; an entry without arguments passes null dummies
            proto = "void %s_body(int ENTRY_SEL_%s)" % (u.name, "".join(", " + q for q in params))
            L.append("%s %s_f(%s)" % (rett, u.name, ", ".join(params) if params else "void"))
            L.append("{ %s; %s_body(0%s); }" % (proto, u.name, "".join(", " + self.cvar(a) for a in u.args)))
            for k, e in enumerate(hosted):
                L.append("void %s_f(void)" % e)
                L.append("{ %s; %s_body(%d%s); }" % (proto, u.name, k + 1, ", 0" * len(u.args)))
            L.append(proto)
        else:
            L.append("%s %s_f(%s)" % (rett, u.name, ", ".join(params) if params else "void"))
        L.append("{")
        # local parameters (not the run-time ones, not the globals)
        for name in u.order:
            s = u.sym[name]
            if s.kind == "param" and not s.is_char and name not in self.param_globals:
                L.append("  const %s %s = %s;" % (CTYPE[s.type], self.cvar(name), self.emit_expr(u, s.value, ctx)))
        # common views
        # character-only blocks (labels, dates) are not storage the translated code touches
        u.commons = [(blk, names) for blk, names in u.commons if not all(u.sym[n].is_char for n in names)]
        for blk, names in u.commons:
            if any(u.sym[n].is_char for n in names):
                raise SyntaxError("COMMON /%s/ mixes character and numeric members" % blk)
            self.register_block(u, blk, names)
        by_block = {}
        for blk, names in u.commons:
            by_block.setdefault(blk, []).extend(names)
        for blk, names in by_block.items():
            L.append("  char *CMB_%s = ref_cm[CM_%s]; long CMO_%s = 0;" % (blk, blk, blk))
            for name in names:
                s = u.sym[name]
                ct = CTYPE[s.type]
                L.append("  %s *%s = (%s*)(CMB_%s + CMO_%s);" % (ct, self.cvar(name), ct, blk, blk))
                cnt = self.emit_dims(u, s, L, ctx)
                L.append("  CMO_%s += %d * (long)(%s);" % (blk, BYTES[s.type], cnt))
            L.append("  (void)CMO_%s;" % blk)
        # dummies with dimensions, locals
        frees = []
        aliases = []
        for name in u.order:
            s = u.sym[name]
            if s.is_char or s.kind in ("param", "common"):
                continue
            if s.kind == "dummy":
                if s.dims:
                    self.emit_dims(u, s, L, ctx)
                continue
            ct = CTYPE[s.type]
            if name in u.equiv:                  # storage-associated with an earlier local: same base address
                own = u.sym[u.equiv[name]]
                if own.kind in ("param", "common", "dummy") or own.type != s.type or not s.dims or not own.dims or own.name in u.equiv:
                    raise SyntaxError("EQUIVALENCE (%s,%s): only same-type local arrays are supported" % (own.name, name))
                self.emit_dims(u, s, L, ctx)
                aliases.append("  %s *%s = %s;" % (ct, self.cvar(name), self.cvar(own.name)))
                continue
            if s.dims:
                cnt = self.emit_dims(u, s, L, ctx)
                if s.save:
                    L.append("  static __thread %s *%s = 0; if (!%s) %s = (%s*)calloc((size_t)(%s), sizeof(%s));" % (
                        ct, self.cvar(name), self.cvar(name), self.cvar(name), ct, cnt, ct))
                else:
                    L.append("  %s *%s = (%s*)calloc((size_t)(%s), sizeof(%s));" % (ct, self.cvar(name), ct, cnt, ct))
                    frees.append(self.cvar(name))
            else:
                if s.save:
                    init = self.emit_expr(u, s.value, ctx) if s.value is not None else "0"
                    L.append("  static __thread %s %s = %s;" % (ct, self.cvar(name), init))
                else:
                    L.append("  %s %s = 0;" % (ct, self.cvar(name)))
        L.extend(aliases)
        for k, e in enumerate(hosted):
            L.append("  if (ENTRY_SEL_ == %d) goto L_entry_%s;" % (k + 1, e))
        # executable part
        for idx in range(body_start, len(u.stmts)):
            n, label, s = u.stmts[idx]
            try:
                self.translate_stmt(u, n, label, s, L, ctx)
            except SyntaxError as ex:
                raise SyntaxError("%s:%d: %s   [%s]" % (u.name, n, ex, s))
        if ctx["do_stack"]:
            raise SyntaxError("%s: unterminated DO" % u.name)
        L.append("L_return: ;")
        for f in frees:
            L.append("  free(%s);" % f)
        if u.kind == "function":
            L.append("  return %s;" % self.cvar(u.name))
        L.append("}")
        # silence unused warnings for locals by construction: compile with -w
        return "\n".join(L)

    def emit_dims(self, u, s, L, ctx):
        """emit the index helpers of array s (name_o, name_s1, ...); returns the C expression of its element count"""
        if not s.dims:
            return "1"
        n = self.cvar(s.name)
        sizes = []
        for k, (lo, hi) in enumerate(s.dims):
            lo_c = self.emit_expr(u, lo, ctx)
            if hi is None:
                sizes.append(None)
            else:
                sizes.append("((%s) - (%s) + 1)" % (self.emit_expr(u, hi, ctx), lo_c))
        # strides
        stride = "1"
        off = []
        for k, (lo, hi) in enumerate(s.dims):
            if k > 0:
                L.append("  const long %s_S%d = %s;" % (n, k, stride))
            lo_c = self.emit_expr(u, lo, ctx)
            off.append("(long)(%s)*(%s)" % (lo_c, stride if k == 0 else "%s_S%d" % (n, k)))
            if sizes[k] is not None:
                stride = "(%s)*(long)%s" % (stride if k == 0 else "%s_S%d" % (n, k), sizes[k])
            elif k != len(s.dims) - 1:
                raise SyntaxError("assumed size in a non-final dimension of %s" % s.name)
        L.append("  const long %s_O = -(%s); (void)%s_O;" % (n, " + ".join(off), n))
        if sizes[-1] is None:
            return "0"
        return stride

    def register_block(self, u, blk, names):
        if blk not in self.blocks:
            self.blocks[blk] = []
            self.block_order.append(blk)
        self.blocks[blk].append((u.name, list(names)))

    def close_dos(self, label, L, ctx):
        while ctx["do_stack"] and ctx["do_stack"][-1] == label:
            ctx["do_stack"].pop()
            L.append("  }}")

    def translate_stmt(self, u, n, label, s, L, ctx):
        if label:
            L.append("L_%s: ;" % label)
        self.translate_simple(u, n, s, L, ctx)
        if label:
            self.close_dos(label, L, ctx)

    # -- unformatted sequential I/O (the restart file of restrt, F:9696-9726 / 9731-9751) ----------------------------------
    UNF_RE = re.compile(r"^(write|read)\s*\(\s*(\d+)\s*\)\s*(.*)$")

    def emit_io_item(self, u, item, L, ctx, ind):
        """one io-list item: a scalar, an array element, a whole array, or an implied DO (item, ..., v = lo, hi)"""
        item = item.strip()
        if item.startswith("(") and match_paren(item, 0) == len(item) - 1:
            parts = split_top(item[1:-1])
            m = re.match(r"^([a-z_][a-z0-9_]*)\s*=\s*(.+)$", parts[-2].strip()) if len(parts) >= 3 else None
            if m:                                        # implied DO
                var, lo, hi = m.group(1), m.group(2), parts[-1]
                cv = self.emit_var(u, var, ctx)
                L.append("%sfor (%s = %s; %s <= %s; %s++) {" % (ind, cv, self.emit_expr(u, parse_expr(lo), ctx), cv,
                                                              self.emit_expr(u, parse_expr(hi), ctx), cv))
                for sub in parts[:-2]:
                    self.emit_io_item(u, sub, L, ctx, ind + "  ")
                L.append("%s}" % ind)
                return
        e = parse_expr(item)
        if e[0] == "var":
            sy = u.sym.get(e[1])
            if sy is None or sy.kind == "param" or sy.is_char:
                raise SyntaxError("io-list item %r" % item)
            if sy.dims:
                L.append("%sref_rec_item((void*)%s, %d * (long)(%s));" % (ind, self.cvar(e[1]), BYTES[sy.type], self.array_count(u, sy, ctx)))
            else:
                L.append("%sref_rec_item(%s, %d);" % (ind, self.emit_actual(u, e, None, 0, ctx), BYTES[sy.type]))
            return
        if e[0] == "call":
            sy = u.sym.get(e[1])
            if sy and sy.dims:
                L.append("%sref_rec_item((void*)&%s[%s], %d);" % (ind, self.cvar(e[1]), self.emit_index(u, sy, e[2], ctx), BYTES[sy.type]))
                return
        raise SyntaxError("io-list item %r is not a variable" % item)

    def translate_unformatted(self, u, s, L, ctx):
        """write(u) list / read(u) list / open(unit=u,...form='unformatted') / close(u): True when the statement was one"""
        m = self.UNF_RE.match(s)
        if m:
            L.append("  ref_rec_begin(%s, %d);" % (m.group(2), 1 if m.group(1) == "write" else 0))
            for item in split_top(m.group(3)):
                if item.strip():
                    self.emit_io_item(u, item, L, ctx, "  ")
            L.append("  ref_rec_end();")
            return True
        m = re.match(r"^open\s*\((.*)\)$", s)
        if m and re.search(r"form\s*=\s*'unformatted'", m.group(1)):
            mu = re.search(r"unit\s*=\s*(\d+)", m.group(1))
            if mu:
                L.append("  ref_unit_open(%s, %d);" % (mu.group(1), 1 if re.search(r"status\s*=\s*'replace'", m.group(1)) else 0))
                self.unf_units.add(mu.group(1))
                return True
        m = re.match(r"^close\s*\(\s*(?:unit\s*=\s*)?(\d+)\s*\)$", s)
        if m and m.group(1) in self.unf_units:
            L.append("  ref_unit_close(%s);" % m.group(1))
            return True
        return False

    def translate_simple(self, u, n, s, L, ctx):
        if IO_RE.match(s) and not re.match(r"^(write|read|open|close|print|format|rewind|backspace|flush)\s*=", s):
            if self.translate_unformatted(u, s, L, ctx):
                return
            L.append("  /* F:%d I/O statement dropped */" % n)
            return
        if s == "continue":
            L.append("  ;")
            return
        if s in ("return",):
            L.append("  goto L_return;")
            return
        if s.startswith("stop"):
            L.append("  ref_stop(%d);" % n)
            return
        if s == "cycle":
            L.append("  continue;")
            return
        if s == "exit":
            L.append("  break;")
            return
        m = re.match(r"^go\s*to\s+(\d+)$", s)
        if m:
            L.append("  goto L_%s;" % m.group(1))
            return
        if re.match(r"^end\s*do$", s):
            if not ctx["do_stack"]:
                raise SyntaxError("END DO without DO")
            ctx["do_stack"].pop()
            L.append("  }}")
            return
        if re.match(r"^end\s*if$", s):
            L.append("  }")
            return
        if s == "else":
            L.append("  } else {")
            return
        m = re.match(r"^else\s*if\s*\(", s)
        if m:
            i0 = s.index("(")
            j = match_paren(s, i0)
            assert s[j + 1:].strip() == "then", s
            L.append("  } else if (%s) {" % self.emit_expr(u, parse_expr(s[i0 + 1:j]), ctx))
            return
        m = re.match(r"^if\s*\(", s)
        if m:
            i0 = s.index("(")
            j = match_paren(s, i0)
            cond = self.emit_expr(u, parse_expr(s[i0 + 1:j]), ctx)
            rest = s[j + 1:].strip()
            if rest == "then":
                L.append("  if (%s) {" % cond)
            else:
                if re.match(r"^\d+\s*,\s*\d+\s*,\s*\d+$", rest):
                    raise SyntaxError("arithmetic IF is not supported")
                L.append("  if (%s) {" % cond)
                self.translate_simple(u, n, rest, L, ctx)
                L.append("  }")
            return
        m = re.match(r"^do\s+while\s*\(", s)
        if m:
            i0 = s.index("(")
            j = match_paren(s, i0)
            ctx["do_stack"].append(None)
            L.append("  {{ while (%s) {" % self.emit_expr(u, parse_expr(s[i0 + 1:j]), ctx))
            # closes with "}}" + one more brace: emit the extra one here
            L[-1] = "  { while (%s) {" % self.emit_expr(u, parse_expr(s[i0 + 1:j]), ctx)
            return
        m = re.match(r"^do\s+(?:(\d+)\s*,?\s*)?([a-z_][a-z0-9_]*)\s*=\s*(.*)$", s)
        if m:
            lab, var, rng = m.group(1), m.group(2), m.group(3)
            parts = split_top(rng)
            if len(parts) not in (2, 3):
                raise SyntaxError("bad DO range")
            lo = self.emit_expr(u, parse_expr(parts[0]), ctx)
            hi = self.emit_expr(u, parse_expr(parts[1]), ctx)
            st = self.emit_expr(u, parse_expr(parts[2]), ctx) if len(parts) == 3 else "1"
            v = self.emit_var(u, var, ctx)
            ctx["tmp"] += 1
            t = ctx["tmp"]
            L.append("  { const int DO_LO%d = %s, DO_HI%d = %s, DO_ST%d = %s; int DO_N%d = (DO_HI%d - DO_LO%d + DO_ST%d) / DO_ST%d;"
                     % (t, lo, t, hi, t, st, t, t, t, t, t))
            L.append("    for (%s = DO_LO%d; DO_N%d > 0; --DO_N%d, %s += DO_ST%d) {" % (v, t, t, t, v, t))
            ctx["do_stack"].append(lab)
            return
        m = re.match(r"^call\s+([a-z_][a-z0-9_]*)\s*(?:\((.*)\))?\s*$", s)
        if m:
            self.emit_call(u, m.group(1), m.group(2), L, ctx, n)
            return
        m = re.match(r"^entry\s+([a-z_][a-z0-9_]*)", s)
        if m:
            L.append("L_entry_%s: ;" % m.group(1))
            u.entries.append(m.group(1))
            return
        # assignment
        eq = find_assign(s)
        if eq < 0:
            raise SyntaxError("unrecognised statement")
        lhs, rhs = s[:eq].strip(), s[eq + 1:].strip()
        le = parse_expr(lhs)
        re_ = parse_expr(rhs)
        if le[0] == "var":
            sy = u.sym.get(le[1])
            if sy is None:
                raise SyntaxError("assignment to undeclared %r" % le[1])
            if sy.dims:      # whole-array fill
                if self.etype(u, re_) == "char":
                    L.append("  /* F:%d character assignment dropped */" % n)
                    return
                cnt = self.array_count(u, sy, ctx)
                L.append("  { long Q_; const %s V_ = %s; for (Q_ = 0; Q_ < (long)(%s); Q_++) %s[Q_] = V_; }" % (
                    CTYPE[sy.type], self.emit_expr(u, re_, ctx), cnt, self.cvar(le[1])))
                return
            if sy.is_char:
                L.append("  /* F:%d character assignment dropped */" % n)
                return
            if sy.kind == "param":
                raise SyntaxError("assignment to a PARAMETER")
        elif le[0] == "call":
            sy = u.sym.get(le[1])
            if sy is None or not sy.dims:
                if sy is not None and sy.is_char:
                    L.append("  /* F:%d character assignment dropped */" % n)
                    return
                raise SyntaxError("statement functions are not supported (%s)" % le[1])
            if sy.is_char:
                L.append("  /* F:%d character assignment dropped */" % n)
                return
        else:
            raise SyntaxError("bad assignment target")
        lt = self.etype(u, le)
        rt = self.etype(u, re_)
        if rt == "char":
            L.append("  /* F:%d character assignment dropped */" % n)
            return
        rc = self.emit_expr(u, re_, ctx)
        if lt != rt:
            rc = "(%s)(%s)" % (CTYPE[lt], rc)
        L.append("  %s = %s;" % (self.emit_expr(u, le, ctx), rc))

    def array_count(self, u, s, ctx):
        out = []
        for lo, hi in s.dims:
            out.append("((long)(%s) - (%s) + 1)" % (self.emit_expr(u, hi, ctx), self.emit_expr(u, lo, ctx)))
        return "*".join(out)

    def emit_call(self, u, name, argstr, L, ctx, n):
        args = [parse_arg(a) for a in split_top(argstr)] if argstr and argstr.strip() else []
        if name.startswith("mpi_"):
            # send/recv buffers go by address, everything else by value (oracle/ref_runtime.c)
            if name == "mpi_allreduce":
                L.append("  ref_mpi_allreduce(%s, %s, %s, %s);" % (
                    self.emit_actual(u, args[0], None, 0, ctx), self.emit_actual(u, args[1], None, 1, ctx),
                    self.emit_expr(u, args[2], ctx), self.emit_expr(u, args[3], ctx)))
            elif name == "mpi_allgather":
                L.append("  ref_mpi_allgather(%s, %s, %s, %s, %s);" % (
                    self.emit_actual(u, args[0], None, 0, ctx), self.emit_expr(u, args[1], ctx),
                    self.emit_actual(u, args[3], None, 3, ctx), self.emit_expr(u, args[4], ctx), self.emit_expr(u, args[2], ctx)))
            elif name == "mpi_barrier":
                L.append("  ref_mpi_barrier();")
            elif name in ("mpi_isend", "mpi_irecv"):      # (buf, count, datatype, peer, tag, comm, request, ierror)
                L.append("  ref_%s(%s, %s, %s, %s, %s);" % (
                    name, self.emit_actual(u, args[0], None, 0, ctx), self.emit_expr(u, args[1], ctx), self.emit_expr(u, args[2], ctx),
                    self.emit_expr(u, args[3], ctx), self.emit_actual(u, args[6], None, 6, ctx)))
            elif name == "mpi_wait":                       # (request, status, ierror)
                L.append("  ref_mpi_wait(%s);" % self.emit_actual(u, args[0], None, 0, ctx))
            else:
                raise SyntaxError("MPI call %s is not supported" % name)
            return
        if name in ("cpu_time", "date_and_time", "flush", "system", "getenv") or name in self.stubs:
            L.append("  /* F:%d call %s dropped */" % (n, name))
            return
        if name not in self.units:
            raise SyntaxError("call of untranslated subroutine %r" % name)
        callee = self.units[name]
        self.externals.add(name)
        if len(args) != len(callee.args):
            raise SyntaxError("call %s: %d actuals for %d dummies" % (name, len(args), len(callee.args)))
        # make sure the callee's dummies are known before types are compared
        self.ensure_decls(callee)
        L.append("  %s_f(%s);" % (name, ", ".join(self.emit_actual(u, a, callee, i, ctx) for i, a in enumerate(args))))

    def ensure_decls(self, cu):
        if getattr(cu, "_decl_done", False):
            return
        cu._decl_done = True
        for idx, (n, label, s) in enumerate(cu.stmts):
            if label is None and self.parse_decl(cu, s):
                continue
            if IO_RE.match(s) and s.startswith("format"):
                continue
            break
        # forget: translate_unit parses again into the same table (idempotent for our subset except commons)
        cu.commons = []

    # -- driver -----------------------------------------------------------------------------------
    def collect_include_params(self, fname):
        """PARAMETERs of an include file become globals: run-time ones are plain variables set by ref_set_params,
        the derived ones are recomputed from them."""
        inc = self.include_stmts(fname)
        tmp = Unit("subroutine", "_inc", [], None, 0)
        for n, label, t in inc:
            self.parse_decl(tmp, t, True)
        for name in tmp.order:
            s = tmp.sym[name]
            if s.kind == "param" and not s.is_char and s.value is not None:
                if s.type is None:
                    s.type = implicit_type(name)
                self.param_globals[name] = (s.type, s.value, tmp)
                self.param_order.append(name)

    def run(self, param_include):
        self.collect_include_params(param_include)
        # transitive closure of the wanted units
        done, todo, bodies = set(), list(self.want), {}
        order = []
        entries = []
        while todo:
            name = todo.pop(0)
            if name in done:
                continue
            if name not in self.units:
                raise SyntaxError("unit %r not found in %s" % (name, self.path))
            if name in self.entry_host:          # translated with (and emitted by) its host unit
                done.add(name)
                entries.append(name)
                if self.entry_host[name] not in done:
                    todo.append(self.entry_host[name])
                continue
            done.add(name)
            before = set(self.externals)
            u = self.units[name]
            u.sym, u.order, u.commons = {}, [], []
            try:
                bodies[name] = self.translate_unit(u)
            except SyntaxError:
                raise
            except Exception as ex:
                raise SyntaxError("%s: internal error %r" % (name, ex))
            order.append(name)
            for x in sorted(self.externals - before):
                if x not in done:
                    todo.append(x)
            for x in sorted(self.externals):
                if x not in done and x not in todo:
                    todo.append(x)
        out = []
        out.append("/* GENERATED by oracle/f03c.py from %s -- derived from GPL-3.0 reference source, do not commit */" % self.path)
        out.append("/* units: %s */" % ", ".join(order))
        # parameter globals
        tmp = None
        for name in self.param_order:
            t, val, tmp = self.param_globals[name]
            out.append("static %s %s;" % (CTYPE[t], self.cvar(name)))
        out.append("static void ref_derive_params(void) {")
        for name in self.param_order:
            t, val, tmp = self.param_globals[name]
            if name in self.runtime_params:
                continue
            out.append("  %s = %s;" % (self.cvar(name), self.emit_expr(tmp, val, {})))
        out.append("}")
        out.append("static void ref_default_params(void) {")
        for name in self.param_order:
            t, val, tmp = self.param_globals[name]
            if name in self.runtime_params:
                out.append("  %s = %s;" % (self.cvar(name), self.emit_expr(tmp, val, {})))
        out.append("}")
        out.append("static int ref_set_param(const char *name, long v) {")
        for name in self.param_order:
            if name in self.runtime_params:
                out.append("  if (!strcmp(name, \"%s\")) { %s = (%s)v; return 0; }" % (name, self.cvar(name), CTYPE[self.param_globals[name][0]]))
        out.append("  return 1;\n}")
        out.append("static long ref_get_param(const char *name) {")
        for name in self.param_order:
            out.append("  if (!strcmp(name, \"%s\")) return (long)%s;" % (name, self.cvar(name)))
        out.append("  return -1;\n}")
        # common blocks
        out.append("enum { %s, CM_COUNT };" % ", ".join("CM_%s" % b for b in self.block_order))
        out.append("static __thread char *ref_cm[CM_COUNT];")
        # prototypes
        for name in order:
            u = self.units[name]
            body = bodies[name]
            out.append(body.split("\n", 1)[0] + ";")
            for e, h in self.entry_host.items():
                if h == name:
                    out.append("void %s_f(void);" % e)
        # block sizes and member lookup: every unit's view, evaluated with the current parameters
        out.append(self.emit_block_tables(order))
        for name in order:
            out.append("")
            out.append("/* ---- %s  (F:%d) ---- */" % (name, self.units[name].line))
            out.append(bodies[name])
        # call table
        out.append("")
        out.append("typedef struct { const char *name; void *fn; int nargs; int rtype; } ref_unit_t;")
        out.append("static const ref_unit_t ref_units[] = {")
        for name in order:
            u = self.units[name]
            out.append("  {\"%s\", (void*)%s_f, %d, %d}," % (name, name, len(u.args), (RANK[u.sym[u.name].type] if u.kind == "function" else 0)))
        for e in self.entry_host:
            if self.entry_host[e] in order:
                out.append("  {\"%s\", (void*)%s_f, 0, 0}," % (e, e))
        out.append("  {0, 0, 0, 0}};")
        return "\n".join(out) + "\n"

    def emit_block_tables(self, order):
        """ref_block_lookup(view_unit or NULL, block, member, &count, &type) -> byte offset; ref_block_bytes(block)"""
        L = []
        L.append("typedef struct { const char *unit, *block, *name; int type; } ref_member_key_t;")
        L.append("static long ref_member(int want_blk, const char *unit, const char *name, long *count, int *type, long *total) {")
        L.append("  long found = -1;")
        for blk in self.block_order:
            L.append("  if (want_blk == CM_%s) {" % blk)
            for uname, names in self.merged_views(blk):
                u = self.units[uname]
                L.append("    if (!unit || !strcmp(unit, \"%s\")) { long off = 0;" % uname)
                ctx = {}
                for name in names:
                    s = u.sym[name]
                    dl = []
                    cnt = self.emit_dims_expr(u, s)
                    L.append("      { long C_ = %s; if (name && !strcmp(name, \"%s\") && found < 0) { found = off; if (count) *count = C_; if (type) *type = %d; }"
                             " off += %d * C_; }" % (cnt, name, RANK[s.type], BYTES[s.type]))
                L.append("      if (total && off > *total) *total = off;")
                L.append("      if (unit) return found; }")
            L.append("  }")
        L.append("  return found;")
        L.append("}")
        L.append("static const char *ref_block_names[] = {%s, 0};" % ", ".join('"%s"' % b for b in self.block_order))
        return "\n".join(L)

    def merged_views(self, blk):
        seen, out = {}, []
        for uname, names in self.blocks[blk]:
            if uname in seen:
                seen[uname].extend(names)
            else:
                seen[uname] = list(names)
                out.append((uname, seen[uname]))
        return out

    def emit_dims_expr(self, u, s):
        if not s.dims:
            return "1"
        parts = []
        for lo, hi in s.dims:
            parts.append("((long)(%s) - (long)(%s) + 1)" % (self.emit_expr(u, hi, {}), self.emit_expr(u, lo, {})))
        return " * ".join(parts)


def parse_arg(a):
    return parse_expr(a)


def implicit_type(name):
    return "int" if name[0] in "ijklmn" else "real"


def find_assign(s):
    """index of the '=' of an assignment statement (not ==, <=, >=, /=), outside parentheses"""
    depth, q = 0, None
    for i, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if i + 1 < len(s) and s[i + 1] == "=":
                return -1
            if i > 0 and s[i - 1] in "<>/=":
                return -1
            return i
    return -1


def translate(path, include_dirs, want, runtime_params=("npc", "mx", "my", "mz", "np0"), param_include="param_080A.h", stubs=()):
    tr = Translator(path, include_dirs, runtime_params, want, stubs)
    return tr.run(param_include), tr


if __name__ == "__main__":
    src, tr = translate(sys.argv[1], [sys.argv[2]], sys.argv[3:])
    sys.stdout.write(src)
