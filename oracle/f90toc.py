#!/usr/bin/env python
"""f90toc.py -- a small Fortran-90-subset to C translator (TEST INFRASTRUCTURE ONLY).

Purpose: no Fortran compiler exists in this image, so the reference cannot be built.  This
script translates the reference's OWN source of the per-tile residual routines
(`/root/reference/src/NKSolver/blockette.F90`: blocketteResCore and every flux / SA /
time-step routine it calls) to C, *reading the source where it lies*; the generated C is
written to `oracle/_ref/` (git-ignored, never committed) and compiled with gcc by
`oracle/Makefile.ref`.  The resulting `oracle/_ref/libblockette_ref.so` is the reference's
arithmetic, statement for statement, and is what pins `oracle/adflow_oracle.c`
(tests/test_oracle_vs_reference.py).

Supported subset (everything the translated routines use): module-level static arrays with
explicit bounds, `use ... only:` renames, scalar/array locals (automatic extents -> VLAs),
optional dummies + present(), do / named do / if-then-else / one-line if / select case,
exit, cycle, return, call (by reference), contained subroutines (-> GCC nested functions),
array-section and whole-array assignments, the intrinsics abs max min sqrt exp log10 sign
dim mod real, `**`, logical/relational operators, kind-suffixed literals.

The module variables of OTHER modules (constants, inputPhysics, flowVarRefState, ...) are not
translated: the hand-written `oracle/ref_env.h` declares them as C globals that the test
harness fills from AdfbParams.
"""
import re
import sys

INTRINSIC_REAL = {"sqrt": "sqrt", "exp": "exp", "log10": "log10", "log": "log", "cos": "cos", "sin": "sin"}


# ----------------------------------------------------------------------------- lexing
def preprocess(text, defined=()):
    """strip comments, join continuations, evaluate #ifdef/#ifndef/#else/#endif, lowercase."""
    out = []
    stack = []
    buf = ""
    for raw in text.split("\n"):
        s = raw.rstrip()
        st = s.strip()
        if st.startswith("#"):
            m = re.match(r"#\s*(ifdef|ifndef|else|endif|if|define|include)\s*(\w+)?", st)
            if not m:
                continue
            d, name = m.group(1), m.group(2)
            if d == "ifdef":
                stack.append(name in defined)
            elif d == "ifndef":
                stack.append(name not in defined)
            elif d == "if":
                stack.append(True)   # `#if PETSC_VERSION_GE(...)`: the current-PETSc branch
            elif d == "else":
                stack[-1] = not stack[-1]
            elif d == "endif":
                stack.pop()
            continue
        if stack and not all(stack):
            continue
        # strip comments (no '!' inside the few string literals of these routines)
        q = None
        cut = len(s)
        for idx, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "!":
                cut = idx
                break
        s = s[:cut].strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            buf += s[:-1].rstrip() + " "
            continue
        line = buf + s
        buf = ""
        # lowercase outside strings
        parts = re.split(r"(\"[^\"]*\"|'[^']*')", line)
        line = "".join(p if (p[:1] in "'\"") else p.lower() for p in parts)
        for piece in split_semicolons(line):
            out.append(piece.strip())
    return out


def split_semicolons(line):
    res, cur, q = [], "", None
    for ch in line:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == ";":
            res.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        res.append(cur)
    return res


TOKEN_RE = re.compile(
    r"\s*(?:(?P<str>\"[^\"]*\"|'[^']*')|(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_[a-z]\w*)?)"
    r"|(?P<dotop>\.(?:and|or|not|eq|ne|lt|le|gt|ge|true|false|eqv|neqv)\.)|(?P<name>[a-z_]\w*)"
    r"|(?P<op>\*\*|==|/=|<=|>=|=>|::|[-+*/(),:%=<>]))")


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m:
            if s[pos:].strip() == "":
                break
            raise SyntaxError("cannot tokenize: %r at %r" % (s, s[pos:pos + 20]))
        pos = m.end()
        for kind in ("str", "num", "dotop", "name", "op"):
            if m.group(kind) is not None:
                toks.append((kind, m.group(kind)))
                break
    return toks


# ----------------------------------------------------------------------------- expression parser
class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val:
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError("expected %r, got %r in %r" % (val, self.peek(), self.t))

    # precedence climbing
    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.peek()[1] in (".or.", ".eqv.", ".neqv."):
            op = self.next()[1]
            a = ("bin", op, a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek()[1] == ".and.":
            self.next()
            a = ("bin", ".and.", a, self.p_not())
        return a

    def p_not(self):
        if self.accept(".not."):
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=", ".eq.": "==", ".ne.": "!=",
           ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}

    def p_rel(self):
        a = self.p_add()
        if self.peek()[1] in self.REL:
            op = self.REL[self.next()[1]]
            a = ("bin", op, a, self.p_add())
        return a

    def p_add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = ("un", op, self.p_mul())
        else:
            a = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.next()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.next()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.accept("**"):
            # right associative; a unary minus in the exponent is allowed
            if self.peek()[1] in ("+", "-"):
                op = self.next()[1]
                b = ("un", op, self.p_pow())
            else:
                b = self.p_pow()
            a = ("pow", a, b)
        return a

    def p_primary(self):
        kind, val = self.next()
        if kind == "num":
            return ("num", val)
        if kind == "str":
            return ("str", val)
        if kind == "dotop" and val in (".true.", ".false."):
            return ("log", val == ".true.")
        if val == "(":
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if kind == "name":
            node = ("name", val)
            while True:
                if self.peek()[1] == "(":
                    self.next()
                    args = []
                    if not self.accept(")"):
                        while True:
                            args.append(self.section_or_expr())
                            if self.accept(")"):
                                break
                            self.expect(",")
                    node = ("call", node, args)
                elif self.peek()[1] == "%":
                    self.next()
                    node = ("member", node, self.next()[1])
                else:
                    break
            return node
        raise SyntaxError("unexpected token %r in %r" % ((kind, val), self.t))

    def section_or_expr(self):
        # handles  a:b , a: , :b , :
        if self.peek()[1] == ":":
            self.next()
            hi = None if self.peek()[1] in (",", ")") else self.expr()
            return ("range", None, hi)
        e = self.expr()
        if self.accept(":"):
            hi = None if self.peek()[1] in (",", ")") else self.expr()
            return ("range", e, hi)
        return e


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.i != len(p.t):
        raise SyntaxError("trailing tokens in %r" % s)
    return e


# ----------------------------------------------------------------------------- symbols
class Array:
    def __init__(self, cname, ctype, bounds, pointer=False, strides=None, base=None):
        self.cname, self.ctype, self.bounds, self.pointer = cname, ctype, bounds, pointer  # bounds: list of (lo_c, ext_c)
        self.strides = strides  # explicit per-dimension strides (Fortran pointer arrays with run-time descriptors)
        # index origin per dimension: the subscript value that maps to offset 0 of `cname`.  Normally the declared
        # lower bound; different for the blockPointers arrays, which ref_env.h stores in one uniform box (0:ib,..)
        # while their DECLARED bounds (needed for `:` sections) are the reference's
        self.base = base

    def origin(self):
        return list(self.base) if self.base else [lo for lo, _ in self.bounds]

    def stride_list(self):
        if self.strides:
            return list(self.strides)
        out, st = [], "1"
        for lo, ext in self.bounds:
            out.append(st)
            st = "%s * (%s)" % (st, ext)
        return out

    @staticmethod
    def descriptor(name, ctype, rank):
        """Fortran pointer array: C pointer + run-time lower bounds, extents, strides"""
        return Array(name, ctype, [("%s_lb[%d]" % (name, d), "%s_n[%d]" % (name, d)) for d in range(rank)],
                     pointer=True, strides=["%s_s[%d]" % (name, d) for d in range(rank)])


class Scope:
    def __init__(self, parent=None):
        self.parent = parent
        self.arrays, self.types, self.renames, self.ptr_scalars, self.consts = {}, {}, {}, set(), {}
        self.cnames = {}  # scalar name -> C identifier when they differ (exported module variables)

    def cname_of(self, n):
        s = self
        while s:
            if n in s.types:
                return s.cnames.get(n, n)
            s = s.parent
        return n

    def lookup_array(self, n):
        s = self
        while s:
            if n in s.arrays:
                return s.arrays[n]
            if n in s.types:
                return None
            s = s.parent
        return None

    def lookup_type(self, n):
        s = self
        while s:
            if n in s.types:
                return s.types[n]
            if n in s.arrays:
                return None
            s = s.parent
        return None

    def is_ptr_scalar(self, n):
        s = self
        while s:
            if n in s.ptr_scalars:
                return True
            if n in s.types or n in s.arrays:
                return False
            s = s.parent
        return False

    def rename(self, n):
        s = self
        while s:
            if n in s.renames:
                return s.renames[n]
            s = s.parent
        return n


class Env:
    """names provided by oracle/ref_env.h: ints (for type inference), arrays, functions."""

    def __init__(self, ints, arrays, int_funcs, subs):
        self.ints, self.arrays, self.int_funcs, self.subs = set(ints), arrays, set(int_funcs), subs


# ----------------------------------------------------------------------------- translator
class Translator:
    def __init__(self, env, rename_modules):
        self.env = env
        self.rename_modules = rename_modules  # module -> prefix for imported names
        self.signatures = {}                  # subroutine -> list of (name, ctype, isarray)
        self.prefix = ""                      # C-name prefix of the current module's procedures
        self.toplevel = set()                 # module procedures (not contained ones) of the current module
        self.all_protos = []                  # prototypes of every translated module procedure
        self.tmp = 0

    # ---- types
    def etype(self, e, sc):
        k = e[0]
        if k == "num":
            v = e[1]
            return "int" if re.fullmatch(r"\d+(_[a-z]\w*)?", v) else "double"
        if k == "log":
            return "int"
        if k == "str":
            return "str"
        if k == "paren":
            return self.etype(e[1], sc)
        if k == "un":
            return "int" if e[1] == ".not." else self.etype(e[2], sc)
        if k == "pow":
            return self.etype(e[1], sc)
        if k == "bin":
            if e[1] in ("==", "!=", "<", "<=", ">", ">=", ".and.", ".or.", ".eqv.", ".neqv."):
                return "int"
            a, b = self.etype(e[2], sc), self.etype(e[3], sc)
            return "int" if a == "int" and b == "int" else "double"
        if k == "name":
            n = e[1]
            t = sc.lookup_type(n)
            if t:
                return t
            arr = sc.lookup_array(n)
            if arr:
                return arr.ctype
            n2 = sc.rename(n)
            if n2 in self.env.ints:
                return "int"
            if n2 in self.env.arrays:
                return self.env.arrays[n2].ctype
            return "double"
        if k == "member":
            return "double"
        if k == "call":
            if e[1][0] == "name":
                n = e[1][1]
                arr = sc.lookup_array(n) or self.env.arrays.get(sc.rename(n))
                if arr:
                    return arr.ctype
                if n in ("max", "min", "abs", "sign", "dim", "mod", "mydim"):
                    return "int" if all(self.etype(a, sc) == "int" for a in e[2]) else "double"
                if n in ("real", "sqrt", "exp", "log10", "log", "cos", "sin"):
                    return "double"
                if n in ("int", "present", "associated") or n in self.env.int_funcs:
                    return "int"
            return "double"
        return "double"

    # ---- expressions
    def num(self, v):
        v = re.sub(r"_[a-z]\w*$", "", v)
        if re.fullmatch(r"\d+", v):
            return v
        v = v.replace("d", "e")
        if "." not in v and "e" not in v:
            return v
        return v

    def index(self, arr, subs, sc, loopmap=None):
        """C index expression of arr(subs); `range` subscripts are replaced through loopmap."""
        if len(subs) != len(arr.bounds):
            raise SyntaxError("rank mismatch for %s: %d subscripts, rank %d" % (arr.cname, len(subs), len(arr.bounds)))
        terms = []
        for lo, s, stride in zip(arr.origin(), subs, arr.stride_list()):
            if s[0] == "range":
                s_c = loopmap.pop(0)
            else:
                s_c = self.ex(s, sc)
            terms.append("((%s) - (%s)) * (%s)" % (s_c, lo, stride))
        idx = " + ".join(terms)
        return "%s[%s]" % (arr.cname, idx)

    def ex(self, e, sc, loopmap=None):
        k = e[0]
        if k == "num":
            return self.num(e[1])
        if k == "log":
            return "1" if e[1] else "0"
        if k == "str":
            return '"%s"' % e[1][1:-1]
        if k == "paren":
            return "(" + self.ex(e[1], sc, loopmap) + ")"
        if k == "un":
            if e[1] == ".not.":
                return "(!(" + self.ex(e[2], sc, loopmap) + "))"
            return "(" + e[1] + self.ex(e[2], sc, loopmap) + ")"
        if k == "pow":
            base = self.ex(e[1], sc, loopmap)
            ex_ = e[2]
            if ex_[0] == "num" and re.fullmatch(r"\d+", re.sub(r"_[a-z]\w*$", "", ex_[1])):
                n = int(re.sub(r"_[a-z]\w*$", "", ex_[1]))
                if self.etype(e[1], sc) == "int":
                    return "f90_ipow(%s, %d)" % (base, n)
                return "f90_powi(%s, %d)" % (base, n)
            return "pow(%s, %s)" % (base, self.ex(ex_, sc, loopmap))
        if k == "bin":
            op = e[1]
            a, b = self.ex(e[2], sc, loopmap), self.ex(e[3], sc, loopmap)
            cop = {".and.": "&&", ".or.": "||", ".eqv.": "==", ".neqv.": "!="}.get(op, op)
            return "(%s %s %s)" % (a, cop, b)
        if k == "name":
            n = e[1]
            arr = sc.lookup_array(n)
            if arr:
                if loopmap is not None and len(loopmap) >= len(arr.bounds):
                    subs = [("range", None, None)] * len(arr.bounds)
                    return self.index(arr, subs, sc, loopmap)
                return arr.cname
            if sc.is_ptr_scalar(n):
                return "(*%s)" % n
            if sc.lookup_type(n):
                return sc.cname_of(n)
            n2 = sc.rename(n)
            if n2 in self.env.arrays:
                arr = self.env.arrays[n2]
                if loopmap is not None and len(loopmap) >= len(arr.bounds):
                    subs = [("range", None, None)] * len(arr.bounds)
                    return self.index(arr, subs, sc, loopmap)
                return arr.cname
            return n2
        if k == "member":
            return "%s.%s" % (self.ex(e[1], sc, loopmap), e[2])
        if k == "call":
            head, args = e[1], e[2]
            if head[0] != "name":
                raise SyntaxError("unsupported call head %r" % (head,))
            n = head[1]
            arr = sc.lookup_array(n) or self.env.arrays.get(sc.rename(n))
            if arr:
                return self.index(arr, args, sc, loopmap)
            a = [self.ex(x, sc, loopmap) for x in args]
            ty = self.etype(e, sc)
            if n in INTRINSIC_REAL:
                return "%s(%s)" % (INTRINSIC_REAL[n], a[0])
            if n == "abs":
                return ("abs(%s)" if ty == "int" else "fabs(%s)") % a[0]
            if n in ("max", "min"):
                fn = ("f90_i%s" if ty == "int" else "f90_d%s") % n
                r = a[0]
                for x in a[1:]:
                    r = "%s(%s, %s)" % (fn, r, x)
                return r
            if n == "sign":
                return "copysign(fabs(%s), %s)" % (a[0], a[1])
            if n in ("dim", "mydim"):
                return "f90_ddim(%s, %s)" % (a[0], a[1])
            if n == "mod":
                return "((%s) %% (%s))" % (a[0], a[1])
            if n == "real":
                return "((double)(%s))" % a[0]
            if n == "int":
                return "((int)(%s))" % a[0]
            if n == "present":
                return "(%s != NULL)" % args[0][1]
            return "%s(%s)" % (sc.rename(n), ", ".join(a))
        if k == "range":
            raise SyntaxError("array section in scalar context")
        raise SyntaxError("cannot translate %r" % (e,))

    # ---- array-section assignment
    def section_dims(self, node, sc):
        """for an lhs/rhs reference return (arr, subs) if it is an array (section) reference else None"""
        if node[0] == "name":
            arr = sc.lookup_array(node[1]) or self.env.arrays.get(sc.rename(node[1]))
            if arr:
                return arr, [("range", None, None)] * len(arr.bounds)
        if node[0] == "call" and node[1][0] == "name":
            arr = sc.lookup_array(node[1][1]) or self.env.arrays.get(sc.rename(node[1][1]))
            if arr and any(s[0] == "range" for s in node[2]):
                return arr, node[2]
        return None

    def assign(self, lhs, rhs, sc, ind):
        sec = self.section_dims(lhs, sc)
        if not sec:
            return [ind + "%s = %s;" % (self.ex(lhs, sc), self.ex(rhs, sc))]
        arr, subs = sec
        # loops over the ranges of the lhs; rhs sections are mapped positionally
        out, loops = [], []
        lhs_idx = []
        for d, s in enumerate(subs):
            lo_decl, ext = arr.bounds[d]
            if s[0] == "range":
                lo = self.ex(s[1], sc) if s[1] is not None else lo_decl
                hi = self.ex(s[2], sc) if s[2] is not None else "(%s) + (%s) - 1" % (lo_decl, ext)
                v = "_s%d" % self.tmp
                self.tmp += 1
                loops.append((v, lo, hi))
                lhs_idx.append(("cexpr", v))
            else:
                lhs_idx.append(s)
        # rhs translation: every array section / whole array gets offsets _v - lo_lhs + lo_rhs
        def rhs_ex(e):
            sec2 = self.section_dims(e, sc)
            if sec2:
                a2, s2 = sec2
                k = 0
                idxs = []
                for d2, ss in enumerate(s2):
                    if ss[0] == "range":
                        v, lo, _hi = loops[k]
                        k += 1
                        lo2 = self.ex(ss[1], sc) if ss[1] is not None else a2.bounds[d2][0]
                        idxs.append(("cexpr", "(%s) - (%s) + (%s)" % (v, lo, lo2)))
                    else:
                        idxs.append(ss)
                return self.index_c(a2, idxs, sc)
            kk = e[0]
            if kk in ("num", "log", "str"):
                return self.ex(e, sc)
            if kk == "paren":
                return "(" + rhs_ex(e[1]) + ")"
            if kk == "un":
                return "(" + ("!" if e[1] == ".not." else e[1]) + rhs_ex(e[2]) + ")"
            if kk == "bin":
                cop = {".and.": "&&", ".or.": "||"}.get(e[1], e[1])
                return "(%s %s %s)" % (rhs_ex(e[2]), cop, rhs_ex(e[3]))
            return self.ex(e, sc)
        body = "%s = %s;" % (self.index_c(arr, lhs_idx, sc), rhs_ex(rhs))
        # innermost loop = first dimension
        for depth, (v, lo, hi) in enumerate(reversed(loops)):
            out.append(ind + "  " * depth + "for (int %s = %s; %s <= %s; %s++) {" % (v, lo, v, hi, v))
        out.append(ind + "  " * len(loops) + body)
        for depth in reversed(range(len(loops))):
            out.append(ind + "  " * depth + "}")
        return out

    def index_c(self, arr, subs, sc):
        terms = []
        for lo, s, stride in zip(arr.origin(), subs, arr.stride_list()):
            s_c = s[1] if s[0] == "cexpr" else self.ex(s, sc)
            terms.append("((%s) - (%s)) * (%s)" % (s_c, lo, stride))
        return "%s[%s]" % (arr.cname, " + ".join(terms))

    # ---- pointer association  p => target(section)
    def associate(self, lhs, rhs, sc, ind):
        P = sc.lookup_array(lhs[1]) or self.env.arrays.get(sc.rename(lhs[1]))
        if not P or not P.strides:
            raise SyntaxError("pointer association to a non-descriptor array %r" % (lhs,))
        p = P.cname
        if rhs[0] == "name":
            T = sc.lookup_array(rhs[1]) or self.env.arrays.get(sc.rename(rhs[1]))
            subs = [("range", None, None)] * len(T.bounds)
            whole = True
        else:
            T = sc.lookup_array(rhs[1][1]) or self.env.arrays.get(sc.rename(rhs[1][1]))
            subs = rhs[2]
            whole = False
        if T is None:
            raise SyntaxError("unknown pointer target %r" % (rhs,))
        out, off, k = [], [], 0
        for d, ((lo, ext), st, sb, org) in enumerate(zip(T.bounds, T.stride_list(), subs, T.origin())):
            if sb[0] == "range":
                lo_r = self.ex(sb[1], sc) if sb[1] is not None else lo
                hi_r = self.ex(sb[2], sc) if sb[2] is not None else "(%s) + (%s) - 1" % (lo, ext)
                # a pointer to a SECTION has lower bound 1; to a whole array it inherits the bounds
                out.append(ind + "%s_lb[%d] = %s; %s_n[%d] = (%s) - (%s) + 1; %s_s[%d] = %s;" % (
                    p, k, lo if whole else "1", p, k, hi_r, lo_r, p, k, st))
                off.append("((%s) - (%s)) * (%s)" % (lo_r, org, st))
                k += 1
            else:
                off.append("((%s) - (%s)) * (%s)" % (self.ex(sb, sc), org, st))
        if k != len(P.bounds):
            raise SyntaxError("rank mismatch in pointer association of %s" % p)
        out.append(ind + "%s = &%s[%s];" % (p, T.cname, " + ".join(off)))
        return out

    # ---- call statements
    def call_stmt(self, name, args, sc):
        name2 = sc.rename(name)
        if name in self.toplevel and name2 == name:
            name2 = self.prefix + name
        sig = self.signatures.get(name) or self.env.subs.get(name2)
        cargs = []
        for q, a in enumerate(args):
            want = sig[q][1] if sig and q < len(sig) else None
            if a[0] == "str":
                cargs.append(self.ex(a, sc))
                continue
            if a[0] == "name":
                n = a[1]
                arr = sc.lookup_array(n) or self.env.arrays.get(sc.rename(n))
                if arr:
                    cargs.append(arr.cname)
                    if sig and q < len(sig) and isinstance(sig[q][2], tuple):
                        cargs.append("(long[]){%s}" % ", ".join(ext for _, ext in arr.bounds))
                        cargs.append("(long[]){%s}" % ", ".join(arr.stride_list()))
                    continue
                if sc.is_ptr_scalar(n):
                    cargs.append(n)
                    continue
                if sc.lookup_type(n) or sc.rename(n) in self.env.ints or True:
                    cargs.append("&" + self.ex(a, sc) if (sc.lookup_type(n) and n not in sc.consts_all()) else self.by_value(a, sc, want))
                    continue
            if a[0] == "call" and a[1][0] == "name" and (sc.lookup_array(a[1][1]) or self.env.arrays.get(sc.rename(a[1][1]))):
                cargs.append("&" + self.ex(a, sc))
                continue
            cargs.append(self.by_value(a, sc, want))
        if sig:
            want_n = sum(3 if isinstance(k, tuple) else 1 for _, _, k in sig)
            cargs += ["NULL"] * (want_n - len(cargs))
        return "%s(%s);" % (name2, ", ".join(cargs))

    def by_value(self, a, sc, want):
        t = want or self.etype(a, sc)
        return "&(%s){%s}" % (t, self.ex(a, sc))


def _consts_all(self):
    s, out = self, set()
    while s:
        out |= set(s.consts)
        s = s.parent
    return out


Scope.consts_all = _consts_all


# ----------------------------------------------------------------------------- declarations
DECL_RE = re.compile(r"^(real|integer|logical|character)\s*(\([^)]*\))?\s*(.*)$")


def split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def parse_decl(line):
    """-> (ctype, attrs(list of str), entities[(name, dimspec or None, init or None)]) or None"""
    m = DECL_RE.match(line)
    if not m or "::" not in line and not re.match(r"^(real|integer|logical)\b", line):
        return None
    base = m.group(1)
    rest = m.group(3)
    if "::" in rest:
        attrs_s, ents_s = rest.split("::", 1)
    else:
        attrs_s, ents_s = "", rest
    attrs = [a for a in split_top(attrs_s.strip().lstrip(","), ",") if a]
    ctype = {"real": "double", "integer": "int", "logical": "int", "character": "char"}[base]
    ents = []
    for ent in split_top(ents_s, ","):
        init = None
        if "=" in ent and "=>" not in ent:
            ent, init = ent.split("=", 1)
        ent = ent.strip()
        mm = re.match(r"^(\w+)\s*(\((.*)\))?$", ent)
        ents.append((mm.group(1), mm.group(3), init.strip() if init else None))
    return ctype, attrs, ents


def dims_from_spec(spec, tr, sc):
    """'0:bbib, 2:max(il,jl)' -> [(lo_c, ext_c)]; ':' or '*' -> None (assumed)"""
    res = []
    for d in split_top(spec, ","):
        d = d.strip()
        if d in (":", "*"):
            res.append(None)
            continue
        parts = split_top(d, ":")
        if len(parts) == 1 and ":" not in d:
            lo, hi = "1", tr.ex(parse_expr(parts[0]), sc)
        else:
            lo_s, hi_s = (d.split(":", 1) + [""])[:2]
            lo = tr.ex(parse_expr(lo_s), sc) if lo_s.strip() else "1"
            if not hi_s.strip():
                res.append(None)
                continue
            hi = tr.ex(parse_expr(hi_s), sc)
        res.append((lo, "(%s) - (%s) + 1" % (hi, lo)))
    return res


# ----------------------------------------------------------------------------- statement translation
class UnitTranslator:
    def __init__(self, tr, lines, module_scope, only=None):
        self.tr, self.lines, self.msc, self.only = tr, lines, module_scope, only
        self.out = []

    def find_subroutines(self):
        """index ranges of top-level subroutines and their contained ones"""
        subs, stack = [], []
        for idx, l in enumerate(self.lines):
            m = re.match(r"^subroutine\s+(\w+)\s*(\((.*)\))?$", l)
            if m:
                stack.append([m.group(1), idx, m.group(3) or "", len(stack)])
            elif re.match(r"^end\s+subroutine", l):
                s = stack.pop()
                subs.append((s[0], s[1], idx, s[2], s[3]))
        return subs

    def collect_signatures(self):
        for name, a, b, args, depth in self.find_subroutines():
            argn = [x.strip() for x in args.split(",") if x.strip()]
            types = {}
            for l in self.lines[a + 1:b]:
                d = parse_decl(l)
                if not d:
                    if re.match(r"^(use|implicit)\b", l):
                        continue
                    if not re.match(r"^(real|integer|logical)", l):
                        break
                    continue
                ctype, attrs, ents = d
                dimattr = next((x for x in attrs if x.startswith("dimension")), None)
                for en, dspec, _ in ents:
                    if en in argn:
                        spec = dspec or (dimattr[dimattr.index("(") + 1:-1] if dimattr else None)
                        kind = False
                        if spec is not None:
                            parts = [x.strip() for x in split_top(spec, ",")]
                            kind = ("desc", len(parts)) if (len(parts) >= 2 and all(x == ":" for x in parts)) else True
                        types[en] = (ctype, kind)
            self.tr.signatures[name] = [(n, types.get(n, ("double", False))[0], types.get(n, ("double", False))[1]) for n in argn]

    def translate(self):
        self.collect_signatures()
        subs = self.find_subroutines()
        top = [s for s in subs if s[4] == 0 and (self.only is None or s[0] in self.only)]
        self.tr.toplevel = set(s[0] for s in subs if s[4] == 0)
        protos = []
        for name, a, b, args, depth in top:
            protos.append(self.proto(name) + ";")
            self.tr.env.subs[self.tr.prefix + name] = self.tr.signatures[name]
        self.tr.all_protos += protos
        self.out += protos + [""]
        for name, a, b, args, depth in top:
            self.out += self.subroutine(name, a, b, self.msc, nested=False)
            self.out.append("")
        return "\n".join(self.out)

    def proto(self, name, nested=False):
        sig = self.tr.signatures[name]
        ps = ", ".join(("%s* %s, long* %s_n, long* %s_s" % (t, n, n, n)) if isinstance(k, tuple) else ("%s* %s" % (t, n))
                       for n, t, k in sig) or "void"
        cname = name if (nested or name not in self.tr.toplevel) else self.tr.prefix + name
        return "%svoid %s(%s)" % ("auto " if nested else "", cname, ps)

    def subroutine(self, name, a, b, parent_scope, nested):
        tr = self.tr
        sc = Scope(parent_scope)
        sig = tr.signatures[name]
        argnames = [n for n, _, _ in sig]
        body = self.lines[a + 1:b]
        # contained subroutines
        contains_at = None
        depth = 0
        for q, l in enumerate(body):
            if re.match(r"^subroutine\b", l):
                depth += 1
            elif re.match(r"^end\s+subroutine", l):
                depth -= 1
            elif l == "contains" and depth == 0:
                contains_at = q
                break
        main = body if contains_at is None else body[:contains_at]
        inner = [] if contains_at is None else body[contains_at + 1:]
        out = [(self.proto(name).replace("void " + name, "void " + name) if not nested else self.proto(name, False)) + " {"]
        decls, code = [], []
        ind = "    "
        # --- declarations
        q = 0
        exec_started = False
        stmts = []
        while q < len(main):
            l = main[q]
            q += 1
            if not exec_started:
                if l.startswith("use "):
                    m = re.match(r"^use\s+(\w+)\s*(,\s*only\s*:\s*(.*))?$", l)
                    mod, only = m.group(1), m.group(3)
                    pref = tr.rename_modules.get(mod)
                    if only:
                        for item in split_top(only, ","):
                            if "=>" in item:
                                loc, ext = [x.strip() for x in item.split("=>")]
                            else:
                                loc = ext = item.strip()
                            if pref is not None:
                                sc.renames[loc] = pref + ext
                            elif loc != ext:
                                sc.renames[loc] = ext
                    elif pref is not None:
                        # whole-module import: every name ref_env.h provides under this prefix
                        for en in list(tr.env.ints) + list(tr.env.arrays) + list(tr.env.subs):
                            if en.startswith(pref):
                                sc.renames[en[len(pref):]] = en
                    continue
                if l.startswith("implicit"):
                    continue
                d = parse_decl(l)
                if d:
                    ctype, attrs, ents = d
                    dimattr = next((x for x in attrs if x.startswith("dimension")), None)
                    is_param = "parameter" in attrs
                    for en, dspec, init in ents:
                        spec = dspec or (dimattr[dimattr.index("(") + 1:-1] if dimattr else None)
                        if en in argnames:
                            kind = next(k for n_, t_, k in sig if n_ == en)
                            if isinstance(kind, tuple):  # assumed shape, rank >= 2: lower bound 1, caller's strides
                                sc.arrays[en] = Array(en, ctype, [("1", "%s_n[%d]" % (en, q_)) for q_ in range(kind[1])],
                                                      pointer=True, strides=["%s_s[%d]" % (en, q_) for q_ in range(kind[1])])
                            elif spec is not None:
                                dims = dims_from_spec(spec, tr, sc)
                                if len(dims) == 1:
                                    lo = dims[0][0] if dims[0] else "1"
                                    sc.arrays[en] = Array(en, ctype, [(lo, "1")], pointer=True)
                                else:
                                    sc.arrays[en] = Array(en, ctype, [(x[0], x[1]) if x else ("1", "1") for x in dims], pointer=True)
                            else:
                                sc.ptr_scalars.add(en)
                                sc.types[en] = ctype
                            continue
                        if spec is not None and "pointer" in attrs:
                            rank = len(split_top(spec, ","))
                            decls.append(ind + "%s* %s = 0; long %s_lb[%d] = {0}, %s_n[%d] = {0}, %s_s[%d] = {0};" % (
                                ctype, en, en, rank, en, rank, en, rank))
                            sc.arrays[en] = Array.descriptor(en, ctype, rank)
                        elif spec is not None:
                            dims = dims_from_spec(spec, tr, sc)
                            total = " * ".join("(%s)" % x[1] for x in dims)
                            decls.append(ind + "%s %s_[%s];" % (ctype, en, total))
                            sc.arrays[en] = Array(en + "_", ctype, dims)
                        else:
                            sc.types[en] = ctype
                            if is_param:
                                sc.consts[en] = True
                                decls.append(ind + "const %s %s = %s;" % (ctype, en, tr.ex(parse_expr(init), sc)))
                            elif init is not None:
                                decls.append(ind + "%s %s = %s;" % (ctype, en, tr.ex(parse_expr(init), sc)))
                            else:
                                decls.append(ind + "%s %s = 0;" % (ctype, en))
                    continue
                exec_started = True
            stmts.append(l)
        out += decls
        for nm, a2, b2, args2, dep2 in self.find_subroutines():
            pass
        # nested prototypes
        inner_subs = []
        if inner:
            sub_ut = UnitTranslator(tr, inner, sc)
            for nm, a2, b2, args2, dep2 in sub_ut.find_subroutines():
                if dep2 == 0:
                    inner_subs.append((nm, a2, b2))
                    out.append(ind + self.proto(nm, True) + ";")
        out += self.block(stmts, sc, ind)
        if inner:
            sub_ut = UnitTranslator(tr, inner, sc)
            for nm, a2, b2 in inner_subs:
                sub_lines = sub_ut.subroutine(nm, a2, b2, sc, nested=True)
                out += [ind + x for x in sub_lines]
        out.append("}")
        return out

    # ---- executable statements
    def block(self, stmts, sc, ind):
        tr = self.tr
        out = []
        sel_stack = []  # (selector_c, first_case_seen)
        for l in stmts:
            # strip construct names:  name: do ...   /  end do name  / else name
            m = re.match(r"^(\w+)\s*:\s*(do|if|select)\b(.*)$", l)
            if m and not re.match(r"^(\w+)\s*:\s*:", l):
                l = m.group(2) + m.group(3)
            try:
                out += self.stmt(l, sc, ind, sel_stack)
            except Exception as ex:
                raise SyntaxError("while translating %r: %s" % (l, ex))
        return out

    def stmt(self, l, sc, ind, sel_stack):
        tr = self.tr
        if re.match(r"^end\s*do\b", l) or re.match(r"^end\s*if\b", l) or l == "endif" or l == "enddo":
            return [ind + "}"]
        if re.match(r"^end\s*select\b", l):
            sel_stack.pop()
            return [ind + "}"]
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", l)
        if m:
            var = m.group(1)
            parts = split_top(m.group(2), ",")
            lo, hi = tr.ex(parse_expr(parts[0]), sc), tr.ex(parse_expr(parts[1]), sc)
            v = tr.ex(("name", var), sc)
            if len(parts) == 3:
                st = tr.ex(parse_expr(parts[2]), sc)
                neg = parts[2].strip().startswith("-")
                cmp_ = ">=" if neg else "<="
                return [ind + "for (%s = %s; %s %s %s; %s += %s) {" % (v, lo, v, cmp_, hi, v, st)]
            return [ind + "for (%s = %s; %s <= %s; %s++) {" % (v, lo, v, hi, v)]
        if l == "do":
            return [ind + "for (;;) {"]
        m = re.match(r"^select\s*case\s*\((.*)\)$", l)
        if m:
            sel_stack.append([tr.ex(parse_expr(m.group(1)), sc), False])
            return [ind + "if (0) {"]
        m = re.match(r"^case\s*\((.*)\)$", l)
        if m:
            sel = sel_stack[-1][0]
            conds = " || ".join("(%s) == (%s)" % (sel, tr.ex(parse_expr(x), sc)) for x in split_top(m.group(1), ","))
            return [ind + "} else if (%s) {" % conds]
        if re.match(r"^case\s+default$", l):
            return [ind + "} else {"]
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then(\s+\w+)?$", l)
        if m:
            return [ind + "} else if (%s) {" % tr.ex(parse_expr(m.group(1)), sc)]
        if re.match(r"^else(\s+\w+)?$", l):
            return [ind + "} else {"]
        if l.startswith("if"):
            # find the matching paren of the condition
            start = l.index("(")
            depth, q = 0, start
            while True:
                if l[q] == "(":
                    depth += 1
                elif l[q] == ")":
                    depth -= 1
                    if depth == 0:
                        break
                q += 1
            cond = tr.ex(parse_expr(l[start + 1:q]), sc)
            rest = l[q + 1:].strip()
            if rest == "then":
                return [ind + "if (%s) {" % cond]
            inner = self.stmt(rest, sc, ind + "    ", sel_stack)
            return [ind + "if (%s) {" % cond] + inner + [ind + "}"]
        if l == "exit":
            return [ind + "break;"]
        if l == "cycle":
            return [ind + "continue;"]
        if l == "return":
            return [ind + "return;"]
        if l == "continue":
            return [ind + ";"]
        if re.match(r"^(print|write)\b", l):  # diagnostics only
            return [ind + "/* %s */;" % l.replace("*/", "* /")]
        if l == "stop" or l.startswith("stop "):
            return [ind + "abort();"]
        m = re.match(r"^allocate\s*\((.*)\)$", l)
        if m:
            node = Parser(tokenize(m.group(1))).expr()
            arr = sc.lookup_array(node[1][1])
            p, out, total, st = arr.cname, [], [], "1"
            for dd, sub in enumerate(node[2]):
                lo = tr.ex(sub[1], sc) if sub[0] == "range" else "1"
                hi = tr.ex(sub[2], sc) if sub[0] == "range" else tr.ex(sub, sc)
                out.append(ind + "%s_lb[%d] = %s; %s_n[%d] = (%s) - (%s) + 1; %s_s[%d] = %s;" % (p, dd, lo, p, dd, hi, lo, p, dd, st))
                st = "%s * %s_n[%d]" % (st, p, dd)
            out.append(ind + "%s = (%s*)malloc(sizeof(%s) * (%s));" % (p, arr.ctype, arr.ctype, st))
            return out
        m = re.match(r"^deallocate\s*\((\w+)\)$", l)
        if m:
            arr = sc.lookup_array(m.group(1))
            return [ind + "free(%s); %s = 0;" % (arr.cname, arr.cname)]
        m = re.match(r"^call\s+(\w+)\s*(\((.*)\))?$", l)
        if m:
            args = []
            if m.group(3) and m.group(3).strip():
                p = Parser(tokenize("f(" + m.group(3) + ")"))
                args = p.expr()[2]
            return [ind + tr.call_stmt(m.group(1), args, sc)]
        # assignment
        toks = tokenize(l)
        depth = 0
        for q, (k, v) in enumerate(toks):
            if v == "(":
                depth += 1
            elif v == ")":
                depth -= 1
            elif v == "=>" and depth == 0:
                return tr.associate(Parser(toks[:q]).expr(), Parser(toks[q + 1:]).expr(), sc, ind)
        depth = 0
        for q, (k, v) in enumerate(toks):
            if v == "(":
                depth += 1
            elif v == ")":
                depth -= 1
            elif v == "=" and depth == 0:
                lhs = Parser(toks[:q]).expr()
                rhs = Parser(toks[q + 1:]).expr()
                return tr.assign(lhs, rhs, sc, ind)
        raise SyntaxError("unrecognised statement")


C_PRELUDE = r"""/* GENERATED by oracle/f90toc.py from %(src)s -- do not edit, do not commit */
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include "ref_env.h"
#include "ref_protos.h"
static inline double f90_dmax(double a, double b) { return a > b ? a : b; }
static inline double f90_dmin(double a, double b) { return a < b ? a : b; }
static inline int f90_imax(int a, int b) { return a > b ? a : b; }
static inline int f90_imin(int a, int b) { return a < b ? a : b; }
static inline double f90_ddim(double a, double b) { return a - b > 0.0 ? a - b : 0.0; }
/* x**n with integer n: the algorithm of libgcc's __powidf2, which is what gfortran calls without
   -ffast-math (binary exponentiation: x**6 = x^2 * x^4) */
static inline double f90_powi(double x, int m) {
    unsigned int n = m < 0 ? -(unsigned int)m : (unsigned int)m;
    double y = n %% 2 ? x : 1.0;
    while (n >>= 1) { x = x * x; if (n %% 2) y *= x; }
    return m < 0 ? 1.0 / y : y;
}
static inline int f90_ipow(int x, int n) { int r = 1; for (int q = 0; q < n; q++) r *= x; return r; }
"""


def translate_module(src_path, only, env, rename_modules, patches=(), defined=(), tr=None, prefix="", module=None):
    text = open(src_path).read()
    lines = preprocess(text, defined)
    if module is not None:   # a file with several modules: keep `module <name>` ... `end module <name>`
        a = next(q for q, l in enumerate(lines) if re.match(r"module\s+%s\s*$" % module, l))
        b = next(q for q, l in enumerate(lines) if q > a and re.match(r"end\s*module", l))
        lines = lines[a:b + 1]
    for patch in patches:
        if patch[0] == "block":  # ("block", first-line regex, last-line regex): drop the lines in between, inclusive
            out, skipping = [], False
            for l in lines:
                if not skipping and re.search(patch[1], l):
                    skipping = True
                if not skipping:
                    out.append(l)
                elif re.search(patch[2], l):
                    skipping = False
            lines = out
        else:
            lines = [re.sub(patch[0], patch[1], l) for l in lines]
    lines = [l for l in lines if l is not None and l.strip() != ""]
    if tr is None:
        tr = Translator(env, rename_modules)
    tr.signatures, tr.prefix, tr.toplevel = {}, prefix, set()
    # module-level declarations
    msc = Scope()
    mdecl = []
    q = 0
    assert lines[0].startswith("module ")
    q = 1
    while q < len(lines) and lines[q] != "contains" and not lines[q].startswith("end module"):
        l = lines[q]
        q += 1
        if l.startswith("use ") or l.startswith("implicit"):
            continue
        d = parse_decl(l)
        if not d:
            continue
        ctype, attrs, ents = d
        dimattr = next((x for x in attrs if x.startswith("dimension")), None)
        is_param = "parameter" in attrs
        for en, dspec, init in ents:
            spec = dspec or (dimattr[dimattr.index("(") + 1:-1] if dimattr else None)
            if spec is not None:
                dims = dims_from_spec(spec, tr, msc)
                if any(x is None for x in dims):
                    if "pointer" in attrs and prefix:  # exported pointer array: global run-time descriptor
                        rank, cn = len(dims), prefix + en
                        mdecl.append("%s* %s = 0; long %s_lb[%d], %s_n[%d], %s_s[%d];" % (ctype, cn, cn, rank, cn, rank, cn, rank))
                        tr.all_protos.append("extern %s* %s; extern long %s_lb[%d], %s_n[%d], %s_s[%d];" % (
                            ctype, cn, cn, rank, cn, rank, cn, rank))
                        tr.env.arrays[cn] = Array.descriptor(cn, ctype, rank)
                        msc.arrays[en] = tr.env.arrays[cn]
                    elif "allocatable" in attrs:  # allocate()/deallocate() manage it at run time
                        rank = len(dims)
                        mdecl.append("static %s* %s = 0; static long %s_lb[%d], %s_n[%d], %s_s[%d];" % (
                            ctype, en, en, rank, en, rank, en, rank))
                        msc.arrays[en] = Array.descriptor(en, ctype, rank)
                    continue  # pointer module array: only usable if ref_env.h provides it
                total = " * ".join("(%s)" % x[1] for x in dims)
                mdecl.append("static %s %s_[%s];" % (ctype, en, total))
                msc.arrays[en] = Array(en + "_", ctype, dims)
            else:
                msc.types[en] = ctype
                if is_param:
                    msc.consts[en] = True
                    mdecl.append("enum { %s = %s };" % (en, tr.ex(parse_expr(init), msc)) if ctype == "int" else
                                 "static const %s %s = %s;" % (ctype, en, tr.ex(parse_expr(init), msc)))
                elif prefix:
                    # module variable of a named module: exported, other modules import it by `use`
                    msc.cnames[en] = prefix + en
                    mdecl.append("%s %s = 0;" % (ctype, prefix + en))
                    tr.all_protos.append("extern %s %s;" % (ctype, prefix + en))
                    if ctype == "int":
                        tr.env.ints.add(prefix + en)
                else:
                    mdecl.append("static %s %s = 0;" % (ctype, en))
    body = lines[q + 1:] if (q < len(lines) and lines[q] == "contains") else []
    # drop the trailing 'end module' (and anything after it: further small modules in the same file)
    for q2, l in enumerate(body):
        if l.startswith("end module"):
            body = body[:q2]
            break
    ut = UnitTranslator(tr, body, msc, only)
    code = ut.translate()
    return C_PRELUDE % {"src": src_path} + "\n".join(mdecl) + "\n\n" + code + "\n", tr


def translate_parameters(src_path, defined=()):
    """C definitions of every scalar `parameter` of a module (used for the reference's constants.F90)."""
    lines = preprocess(open(src_path).read(), defined)
    tr = Translator(Env([], {}, [], {}), {})
    sc = Scope()
    out = ["/* GENERATED by oracle/f90toc.py from %s -- do not edit, do not commit */" % src_path]
    for l in lines:
        d = parse_decl(l)
        if not d:
            continue
        ctype, attrs, ents = d
        if "parameter" not in attrs or ctype == "char":
            continue
        for en, dspec, init in ents:
            if dspec is not None or init is None:
                continue
            sc.types[en] = ctype
            val = tr.ex(parse_expr(init), sc)
            out.append("enum { %s = %s };" % (en, val) if ctype == "int" else "static const double %s = %s;" % (en, val))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    print(__doc__)
    sys.exit(0)
