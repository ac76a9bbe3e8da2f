"""fortran_subset.py -- a small interpreter for the subset of Fortran 90 in which lpm-v2 writes its direct-sum hot path.

TEST INFRASTRUCTURE ONLY (as everything under oracle/): nothing in the product path imports this file.

Why it exists.  The image has no Fortran compiler, so the reference cannot be built, and it ships no golden vectors for
the BVE / planar / beta-plane sums.  oracle/lpm_oracle.c restates those loops by hand.  This interpreter removes the
"by hand" from the check of that restatement: it reads the reference's OWN source text from /root/reference at
fixture-generation time, parses the procedures on the path (BVESphereVelocity, timestepPrivate, SetStreamFunctionsOnMesh,
LoadBalance, the PSE operators, the PolyMesh2d mesh builder with its seed files, ...) and executes them statement by statement.  oracle/make_refsrc_fixtures.py stores inputs and outputs
under tests/golden/refsrc_*.npz, and tests/test_refsrc_golden.py demands that the C oracle reproduces them BIT FOR BIT.
What is interpreted is the reference's text; what this file supplies is only the language: operator precedence,
left-to-right evaluation of operators of equal precedence, integer division, do / if, array sections and whole-array
assignment, argument association, and the libm intrinsics (Python's math module calls the same glibc functions a
gfortran build links).  It is not a compiler: a gfortran build may still differ from these vectors where the compiler
contracts a*b+c into an FMA (-march=native) or where libm versions differ; on baseline x86-64 without -ffast-math
gfortran evaluates these expressions as written (parentheses are protected, no reassociation) -- and that is how the
reference builds itself: CMakeLists.txt:27 sets "-O3 -fopenmp -ffree-form -cpp" for gfortran, no -march, no -ffast-math.

Semantics implemented
  values        int, float (IEEE double; Python floats), bool, FArr (rank-1 array view: storage list, offset, length,
                lower bound), Obj (derived type: a dict of components).  Identifiers are case-insensitive.
  expressions   .or. < .and. < .not. < relational < + - (binary and unary) < * / < ** ; equal precedence associates to
                the left (** to the right); integer / integer truncates toward zero; x ** n with an integer literal n
                is repeated multiplication (what gfortran emits for small n); array operands work element by element;
                division by zero, log(0), overflow give the IEEE result instead of a Python exception.
  sum(a)        elements added in index order starting from the first (gfortran's inline expansion).
  statements    assignment (scalar, element, section, whole array, component), do / enddo (with step; cycle, exit), if /
                else if / else / endif, one-line if, select case, call, return, allocate; rank-2 arrays as far as the path
                needs them (elements, column sections a(lo:hi, j), whole-array assignment and arithmetic,
                MATMUL(matrix, vector) in gfortran's inline order); open / list-directed read / close on the reference's
                own data files (the mesh seeds); write / print / deallocate are no-ops; procedure dummy arguments;
                declarations only allocate local arrays of constant size
                (dimension(3) :: xi) and evaluate parameter constants; "!$acc" lines are comments; calls to MPI_BCAST,
                LogMessage and the like are no-ops (one rank: numProcs = 1, procRank = 0 -- a broadcast to oneself).
  procedures    looked up by name in the files given to Program(); generic interfaces pick the specific procedure by
                argument count and, where that is ambiguous (New), by the derived type of the first argument; dummy
                arguments are associated by reference for arrays and derived types and copied in / out for scalars.
"""
import math
import os
import re

REFERENCE_ROOT = os.environ.get("LPM_REFERENCE_ROOT", "/root/reference")


class FortranError(Exception):
    pass


class Obj:
    """A derived-type value: components by (lower-case) name."""

    def __init__(self, _type=None, **kw):
        self.__dict__["f"] = {k.lower(): v for k, v in kw.items()}
        self.__dict__["tname"] = None if _type is None else _type.lower()     # derived-type name: picks the specific
                                                                              # procedure of a generic like New(...)

    def __getattr__(self, k):
        try:
            return self.__dict__["f"][k.lower()]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self.__dict__["f"][k.lower()] = v


class FArr:
    """Rank-1 array (or a section of one): element i lives at v[off + i - lb]."""
    __slots__ = ("v", "off", "n", "lb")

    def __init__(self, v, off=0, n=None, lb=1):
        self.v, self.off, self.lb = v, off, lb
        self.n = len(v) - off if n is None else n

    @staticmethod
    def zeros(n, lb=1, value=0.0):
        return FArr([value] * n, 0, n, lb)

    @staticmethod
    def of(seq, lb=1, kind=float):
        return FArr([kind(s) for s in seq], 0, None, lb)

    def get(self, i):
        k = i - self.lb
        if k < 0 or k >= self.n:
            raise FortranError(f"index {i} outside [{self.lb}, {self.lb + self.n - 1}]")
        return self.v[self.off + k]

    def set(self, i, x):
        k = i - self.lb
        if k < 0 or k >= self.n:
            raise FortranError(f"index {i} outside [{self.lb}, {self.lb + self.n - 1}]")
        self.v[self.off + k] = x

    def section(self, lo, hi):
        lo = self.lb if lo is None else lo
        hi = self.lb + self.n - 1 if hi is None else hi
        if hi < lo:
            return FArr(self.v, self.off, 0, 1)
        if lo < self.lb or hi > self.lb + self.n - 1:
            raise FortranError(f"section {lo}:{hi} outside [{self.lb}, {self.lb + self.n - 1}]")
        return FArr(self.v, self.off + lo - self.lb, hi - lo + 1, 1)

    def tolist(self):
        return self.v[self.off:self.off + self.n]

    def assign(self, val):
        if isinstance(val, FArr):
            if val.n != self.n:
                raise FortranError(f"array assignment of {val.n} elements to {self.n}")
            self.v[self.off:self.off + self.n] = val.tolist()       # the right-hand side is evaluated first
        else:
            self.v[self.off:self.off + self.n] = [val] * self.n


class FMat:
    """Rank-2 array with constant bounds 1..n1, 1..n2 (column-major), as far as the path needs one: the 3 x 3 projection
    of the sphere operators (element access, whole-array assignment, MATMUL with a vector)."""
    __slots__ = ("v", "n1", "n2")

    def __init__(self, n1, n2, v=None):
        self.n1, self.n2 = n1, n2
        self.v = [0.0] * (n1 * n2) if v is None else v

    def _k(self, i, j):
        if not (1 <= i <= self.n1 and 1 <= j <= self.n2):
            raise FortranError(f"index ({i}, {j}) outside a {self.n1} x {self.n2} array")
        return (i - 1) + (j - 1) * self.n1

    def get(self, i, j):
        return self.v[self._k(i, j)]

    def set(self, i, j, x):
        self.v[self._k(i, j)] = x

    def assign(self, val):
        if isinstance(val, FMat):
            if (val.n1, val.n2) != (self.n1, self.n2):
                raise FortranError("array assignment of another shape")
            self.v[:] = list(val.v)
        else:
            self.v[:] = [val] * len(self.v)

    def column(self, lo, hi, j):
        """a(lo:hi, j) as a rank-1 view (contiguous in column-major storage)"""
        lo = 1 if lo is None else lo
        hi = self.n1 if hi is None else hi
        if not (1 <= lo and hi <= self.n1 and 1 <= j <= self.n2):
            raise FortranError(f"section ({lo}:{hi}, {j}) outside a {self.n1} x {self.n2} array")
        return FArr(self.v, (j - 1) * self.n1 + lo - 1, max(hi - lo + 1, 0), 1)


def _matmul(a, b):
    """MATMUL(matrix, vector) as gfortran expands it inline: c = 0; do k; do i; c(i) = c(i) + a(i,k) * b(k)."""
    if not isinstance(a, FMat) or not isinstance(b, FArr) or a.n2 != b.n:
        raise FortranError("matmul: only matrix x vector is supported")
    bv, c = b.tolist(), [0.0] * a.n1
    for k in range(a.n2):
        for i in range(a.n1):
            c[i] = c[i] + a.v[i + k * a.n1] * bv[k]
    return FArr(c)


# ---- source text -> logical lines ------------------------------------------------------------------------------------
def _strip_comment(line):
    q = None
    for k, c in enumerate(line):
        if q:
            if c == q:
                q = None
        elif c in "'\"":
            q = c
        elif c == "!":
            return line[:k]
    return line


def _lower_outside_strings(line):
    out, q = [], None
    for c in line:
        if q:
            out.append(c)
            if c == q:
                q = None
        else:
            if c in "'\"":
                q = c
            out.append(c.lower())
    return "".join(out)


def logical_lines(text):
    """Comment-free, continuation-joined, lower-cased statements with their first source line number."""
    out, cur, start = [], "", None
    for no, raw in enumerate(text.splitlines(), 1):
        line = _strip_comment(raw.replace("\t", " ")).strip()
        if not line:
            continue
        if cur:
            if line.startswith("&"):
                line = line[1:].lstrip()
        else:
            start = no
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        cur += line
        for part in cur.split(";") if "'" not in cur and '"' not in cur else [cur]:
            if part.strip():
                out.append((start, _lower_outside_strings(part.strip())))
        cur = ""
    return out


# ---- expressions -----------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    (?P<num>(\d+\.\d*|\.\d+|\d+)([ed][+-]?\d+)?(_\w+)?)
  | (?P<dot>\.(and|or|not|true|false|eq|ne|lt|le|gt|ge|eqv|neqv)\.)
  | (?P<name>[a-z_]\w*)
  | (?P<str>'[^']*'|"[^"]*")
  | (?P<op>\*\*|==|/=|<=|>=|=>|\(/|/\)|[-+*/(),:%=<>\[\]])
  | (?P<ws>\s+)
""", re.X)
_REL = {"==": "==", ".eq.": "==", "/=": "/=", ".ne.": "/=", "<": "<", ".lt.": "<", "<=": "<=", ".le.": "<=",
        ">": ">", ".gt.": ">", ">=": ">=", ".ge.": ">="}


def tokenize(s):
    toks, k = [], 0
    while k < len(s):
        m = _TOKEN.match(s, k)
        if not m:
            raise FortranError(f"cannot tokenize {s[k:k + 20]!r} in {s!r}")
        k = m.end()
        if m.lastgroup == "ws":
            continue
        t = m.group(m.lastgroup)
        if m.lastgroup == "num":
            body = re.sub(r"_\w+$", "", t)
            is_real = ("." in body or "e" in body or "d" in body or bool(re.search(r"_k?real|_dp|_8$", t))) and not t.endswith("_kint")
            toks.append(("num", float(body.replace("d", "e")) if is_real else int(body)))
        elif m.lastgroup == "dot":
            toks.append(("op", t))
        elif m.lastgroup == "name":
            toks.append(("name", t))
        elif m.lastgroup == "str":
            toks.append(("str", t[1:-1]))
        else:
            toks.append(("op", t))
    return toks


class Parser:
    def __init__(self, toks):
        self.t, self.k = toks, 0

    def peek(self, v=None):
        if self.k >= len(self.t):
            return None if v is None else False
        return self.t[self.k] if v is None else self.t[self.k] == ("op", v)

    def take(self, v=None):
        tok = self.t[self.k] if self.k < len(self.t) else None
        if tok is None or (v is not None and tok != ("op", v)):
            raise FortranError(f"expected {v!r}, found {tok!r}")
        self.k += 1
        return tok

    def done(self):
        return self.k >= len(self.t)

    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.peek(".or."):
            self.take()
            a = ("or", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek(".and."):
            self.take()
            a = ("and", a, self.p_not())
        return a

    def p_not(self):
        if self.peek(".not."):
            self.take()
            return ("not", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_add()
        tok = self.peek()
        if tok and tok[0] == "op" and tok[1] in _REL:
            self.take()
            return ("rel", _REL[tok[1]], a, self.p_add())
        return a

    def p_add(self):
        if self.peek("-"):
            self.take()
            a = ("neg", self.p_mul())
        elif self.peek("+"):
            self.take()
            a = self.p_mul()
        else:
            a = self.p_mul()
        while self.peek("+") or self.peek("-"):
            op = self.take()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek("*") or self.peek("/"):
            op = self.take()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek("**"):
            self.take()
            if self.peek("-"):              # a ** -b
                self.take()
                return ("pow", a, ("neg", self.p_pow()))
            return ("pow", a, self.p_pow())
        return a

    def p_args(self):
        args = []
        self.take("(")
        if self.peek(")"):
            self.take()
            return args
        while True:
            lo = None
            if not self.peek(":"):
                lo = self.expr()
            if self.peek(":"):
                self.take()
                hi = None if (self.peek(",") or self.peek(")")) else self.expr()
                args.append(("sec", lo, hi))
            elif self.peek("="):            # keyword argument: name = value
                self.take()
                args.append(("kw", lo, self.expr()))
            else:
                args.append(lo)
            if self.peek(","):
                self.take()
                continue
            self.take(")")
            return args

    def p_primary(self):
        tok = self.peek()
        if tok is None:
            raise FortranError("unexpected end of expression")
        if tok[0] == "num":
            self.take()
            return ("lit", tok[1])
        if tok[0] == "str":
            self.take()
            return ("lit", tok[1])
        if tok == ("op", ".true."):
            self.take()
            return ("lit", True)
        if tok == ("op", ".false."):
            self.take()
            return ("lit", False)
        if tok == ("op", "("):
            self.take()
            e = self.expr()
            self.take(")")
            return ("paren", e)
        if tok in (("op", "["), ("op", "(/")):      # array constructor
            close = "]" if tok[1] == "[" else "/)"
            self.take()
            items = []
            while not self.peek(close):
                items.append(self.expr())
                if self.peek(","):
                    self.take()
            self.take(close)
            return ("array", items)
        if tok[0] == "name":
            return self.p_ref()
        raise FortranError(f"unexpected token {tok!r}")

    def p_ref(self):
        parts = []
        while True:
            name = self.take()[1]
            args = self.p_args() if self.peek("(") else None
            parts.append((name, args))
            if self.peek("%"):
                self.take()
                continue
            return ("ref", parts)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.done():
        raise FortranError(f"trailing tokens in {s!r}")
    return e


# ---- statements ------------------------------------------------------------------------------------------------------
_DECL = re.compile(r"^(real|integer|logical|type|class|character|double\s+precision|complex)\b")
_PROC = re.compile(r"^(?:(?:pure|recursive|elemental|impure)\s+)*(?:(?:real|integer|logical)\s*(?:\([^)]*\))?\s+)?(subroutine|function)\s+(\w+)\s*(?:\(([^)]*)\))?\s*(?:result\s*\(\s*(\w+)\s*\))?\s*$")
IGNORED_CALLS = {"mpi_bcast", "mpi_barrier", "mpi_allreduce", "mpi_reduce", "logmessage", "initlogger", "starttimer", "endtimer"}


def _split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
        if c == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += c
    if cur.strip():
        out.append(cur.strip())
    return out


def _match_paren(s, k):
    depth = 0
    for j in range(k, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise FortranError(f"unbalanced parentheses in {s!r}")


class Proc:
    def __init__(self, kind, name, args, result, body, decls, file, line):
        self.kind, self.name, self.args, self.result = kind, name, args, result
        self.body, self.decls, self.file, self.line = body, decls, file, line


def _parse_decl(stmt):
    """-> list of (name, dims-source or None, init-source or None, is_parameter)"""
    if "::" in stmt:
        head, ents = stmt.split("::", 1)
    else:
        return []                                   # old-style declarations are not used on the path
    dim = re.search(r"dimension\s*\(", head)
    hdims = None
    if dim:
        j = _match_paren(head, dim.end() - 1)
        hdims = head[dim.end():j]
    is_param = bool(re.search(r"\bparameter\b", head))
    base = re.sub(r"\s+", "", re.match(r"^(double\s+precision|\w+\s*(\([^)]*\))?)", head).group(1))   # real(kreal), type(faces), ...
    out = []
    for ent in _split_top(ents):
        init = None
        if "=" in ent and "=>" not in ent:
            ent, init = [p.strip() for p in ent.split("=", 1)]
        m = re.match(r"^(\w+)\s*(?:\((.*)\))?$", ent.strip())
        if not m:
            continue
        out.append((m.group(1), m.group(2) if m.group(2) is not None else hdims, init, is_param, base))
    return out


def parse_block(lines, k, enders):
    """Statements from lines[k:] up to (not including) one whose first word(s) are in `enders` -> (stmts, k, ender)"""
    stmts = []
    while k < len(lines):
        no, s = lines[k]
        word = (re.match(r"^(end\s*(?:do|if|subroutine|function|module|program|type|interface|select|where)?)(?:\s+\w+)?\s*$", s)
                or re.match(r"^(else\s*if|elseif)\s*\(", s) or re.match(r"^(else)\s*$", s)
                or re.match(r"^(case)\s*(\(|default)", s))
        if word:
            w = re.sub(r"\s+", "", word.group(1))
            if w in enders or (w.startswith("end") and "end" in enders):
                return stmts, k, w
        if _DECL.match(s) and "function" not in s.split("::")[0]:
            stmts.append(("decl", no, _parse_decl(s)))
            k += 1
            continue
        if re.match(r"^(use|implicit|private|public|save|external|intrinsic|interface|module\s+procedure|procedure|contains|data|format|\d+\s+format)\b", s):
            k += 1
            continue
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s)
        if m:
            parts = _split_top(m.group(2))
            body, k2, _ = parse_block(lines, k + 1, {"enddo"})
            stmts.append(("do", no, m.group(1), parse_expr(parts[0]), parse_expr(parts[1]),
                          parse_expr(parts[2]) if len(parts) > 2 else None, body))
            k = k2 + 1
            continue
        m = re.match(r"^do\s+while\s*\((.*)\)\s*$", s)
        if m:
            body, k2, _ = parse_block(lines, k + 1, {"enddo"})
            stmts.append(("while", no, parse_expr(m.group(1)), body))
            k = k2 + 1
            continue
        if re.match(r"^if\s*\(", s):
            j = _match_paren(s, s.index("("))
            cond, rest = parse_expr(s[s.index("(") + 1:j]), s[j + 1:].strip()
            if rest == "then":
                arms, other = [], None
                body, k2, end = parse_block(lines, k + 1, {"elseif", "else", "endif"})
                arms.append((cond, body))
                while end in ("elseif", "else"):
                    no2, s2 = lines[k2]
                    if end == "elseif":
                        j2 = _match_paren(s2, s2.index("("))
                        c2 = parse_expr(s2[s2.index("(") + 1:j2])
                        body, k2, end = parse_block(lines, k2 + 1, {"elseif", "else", "endif"})
                        arms.append((c2, body))
                    else:
                        other, k2, end = parse_block(lines, k2 + 1, {"endif"})
                stmts.append(("if", no, arms, other))
                k = k2 + 1
            else:
                inner, _, _ = parse_block([(no, rest)], 0, set())
                stmts.append(("if", no, [(cond, inner)], None))
                k += 1
            continue
        m = re.match(r"^select\s*case\s*\((.*)\)\s*$", s)
        if m:
            sel, arms, default = parse_expr(m.group(1)), [], None
            _, k2, end = parse_block(lines, k + 1, {"case", "endselect"})
            while end == "case":
                head = lines[k2][1]
                body, k3, end = parse_block(lines, k2 + 1, {"case", "endselect"})
                if re.match(r"^case\s*default", head):
                    default = body
                else:
                    vals = head[head.index("(") + 1:_match_paren(head, head.index("("))]
                    arms.append(([parse_expr(v) for v in _split_top(vals)], body))
                k2 = k3
            stmts.append(("select", no, sel, arms, default))
            k = k2 + 1
            continue
        m = re.match(r"^(open|close|read)\s*\((.*)$", s)
        if m:
            j = _match_paren(s, s.index("("))
            ctl = Parser(tokenize(s[s.index("("):j + 1].replace("*", "0"))).p_args()      # read(unit, *): the format is not used
            items = [parse_expr(x) for x in _split_top(s[j + 1:])] if m.group(1) == "read" else []
            stmts.append((m.group(1), no, ctl, items))
            k += 1
            continue
        if re.match(r"^(write|print|deallocate|nullify)\b", s):      # logging and clean-up: nothing to do
            k += 1
            continue
        m = re.match(r"^call\s+(\w+)\s*(\(.*\))?\s*$", s)
        if m:
            args = Parser(tokenize(m.group(2))).p_args() if m.group(2) and m.group(1) not in IGNORED_CALLS else []
            stmts.append(("call", no, m.group(1), args))
            k += 1
            continue
        if s in ("return", "continue", "cycle", "exit"):
            stmts.append((s, no))
            k += 1
            continue
        m = re.match(r"^allocate\s*\((.*)\)\s*$", s)
        if m:
            for item in _split_top(m.group(1)):
                if not re.match(r"^stat\s*=", item):
                    stmts.append(("allocate", no, parse_expr(item)))
            k += 1
            continue
        if re.match(r"^(stop|rewind|backspace|inquire)\b", s):
            stmts.append(("unsupported", no, s))
            k += 1
            continue
        # assignment: the top-level '=' that is not part of ==, /=, <=, >=, =>
        depth, eq = 0, -1
        for j, c in enumerate(s):
            if c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
            elif c == "=" and depth == 0 and s[j - 1] not in "=/<>" and s[j + 1:j + 2] not in ("=", ">"):
                eq = j
                break
        if eq < 0:
            raise FortranError(f"line {no}: cannot parse statement {s!r}")
        stmts.append(("assign", no, parse_expr(s[:eq]), parse_expr(s[eq + 1:])))
        k += 1
    return stmts, k, None


def parse_file(path):
    """-> ({procedure name: Proc}, [module-level declarations with initialisers], {generic name: [specific names]})"""
    lines = logical_lines(open(path, errors="replace").read())
    procs, module_decls, generics, k = {}, [], {}, 0
    while k < len(lines):
        no, s = lines[k]
        g = re.match(r"^interface\s+(\w+)\s*$", s)
        if g:
            k += 1
            while k < len(lines) and not re.match(r"^end\s*interface", lines[k][1]):
                mp = re.match(r"^module\s+procedure\s+(.*)$", lines[k][1])
                if mp:
                    generics.setdefault(g.group(1), []).extend(n.strip() for n in mp.group(1).split(","))
                k += 1
            k += 1
            continue
        m = _PROC.match(s)
        if m and not s.startswith("end"):
            kind, name = m.group(1), m.group(2)
            args = [a.strip() for a in (m.group(3) or "").split(",") if a.strip()]
            k2 = k + 1
            while k2 < len(lines) and not re.match(r"^end(\s*(subroutine|function)(\s+\w+)?)?\s*$", lines[k2][1]):
                k2 += 1
            # parsed on first use: most procedures of these files are off the path and use statements outside the subset
            procs.setdefault(name, Proc(kind, name, args, m.group(4) or name, lines[k + 1:k2], None, path, no))
            k = k2 + 1
            continue
        if _DECL.match(s) and "::" in s and "=" in s.split("::", 1)[1]:
            module_decls.append(("decl", no, _parse_decl(s)))       # parameters and initialised module variables
        k += 1
    return procs, module_decls, generics


# ---- execution -------------------------------------------------------------------------------------------------------
class _Return(Exception):
    pass


class _Cycle(Exception):
    pass


class _Exit(Exception):
    pass


def _zero_of(base):
    return 0 if base.startswith("integer") else False if base.startswith("logical") else "" if base.startswith("character") else 0.0


def _idiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _fdiv(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return float("nan")
        return math.copysign(float("inf"), a) * math.copysign(1.0, b)


def _safe(fn, overflow=float("inf")):
    def g(x):
        try:
            return fn(x)
        except OverflowError:
            return overflow if not callable(overflow) else overflow(x)
        except ValueError:
            return float("nan")
    return g


def _log(x):
    if x == 0:
        return float("-inf")
    if x < 0 or x != x:
        return float("nan")
    return math.log(x)


def _elementwise(fn, *a):
    mats = [x for x in a if isinstance(x, FMat)]
    if mats:
        m = mats[0]
        if any((x.n1, x.n2) != (m.n1, m.n2) for x in mats) or any(isinstance(x, FArr) for x in a):
            raise FortranError("array operands of different shapes")
        cols = [x.v if isinstance(x, FMat) else [x] * len(m.v) for x in a]
        return FMat(m.n1, m.n2, [fn(*row) for row in zip(*cols)])
    arrs = [x for x in a if isinstance(x, FArr)]
    if not arrs:
        return fn(*a)
    n = arrs[0].n
    if any(x.n != n for x in arrs):
        raise FortranError("array operands of different sizes")
    cols = [x.tolist() if isinstance(x, FArr) else [x] * n for x in a]
    return FArr([fn(*row) for row in zip(*cols)], 0, n, 1)


def _sum(a, mask=None):
    v = a.tolist()
    if mask is not None:
        v = [x for x, m in zip(v, mask.tolist()) if m]
    if not v:
        return 0.0
    s = v[0]
    for x in v[1:]:
        s = s + x
    return s


INTRINSICS = {
    "sqrt": lambda x: _elementwise(_safe(math.sqrt), x), "dsqrt": lambda x: _elementwise(_safe(math.sqrt), x),
    "exp": lambda x: _elementwise(_safe(math.exp), x), "dexp": lambda x: _elementwise(_safe(math.exp), x),
    "log": lambda x: _elementwise(_log, x), "dlog": lambda x: _elementwise(_log, x),
    "cos": lambda x: _elementwise(_safe(math.cos), x), "sin": lambda x: _elementwise(_safe(math.sin), x),
    "dcos": lambda x: _elementwise(_safe(math.cos), x), "dsin": lambda x: _elementwise(_safe(math.sin), x),
    "tan": lambda x: _elementwise(_safe(math.tan), x), "acos": lambda x: _elementwise(_safe(math.acos), x),
    "asin": lambda x: _elementwise(_safe(math.asin), x), "atan": lambda x: _elementwise(_safe(math.atan), x),
    "cosh": lambda x: _elementwise(_safe(math.cosh), x),
    "sinh": lambda x: _elementwise(_safe(math.sinh, lambda v: math.copysign(float("inf"), v)), x),
    "tanh": lambda x: _elementwise(math.tanh, x),
    "atan2": lambda y, x: _elementwise(math.atan2, y, x), "datan2": lambda y, x: _elementwise(math.atan2, y, x),
    "abs": lambda x: _elementwise(abs, x), "dabs": lambda x: _elementwise(abs, x),
    "real": lambda x, kind=None: _elementwise(float, x), "dble": lambda x: _elementwise(float, x),
    "int": lambda x, kind=None: _elementwise(int, x),
    "min": lambda *a: _elementwise(min, *a), "max": lambda *a: _elementwise(max, *a),
    "mod": lambda a, b: _elementwise(lambda p, q: math.fmod(p, q) if isinstance(p, float) or isinstance(q, float) else p - _idiv(p, q) * q, a, b),
    "sum": _sum, "size": lambda a, dim=None: (a.n1 if dim == 1 else a.n2 if dim == 2 else a.n1 * a.n2) if isinstance(a, FMat) else a.n, "allocated": lambda a: a is not None, "associated": lambda a: a is not None,
    "dot_product": lambda a, b: _sum(_elementwise(lambda p, q: p * q, a, b)),
    "maxval": lambda a: max(a.tolist()), "minval": lambda a: min(a.tolist()), "matmul": _matmul,
    "trim": lambda a: a, "present": lambda a: a is not None,
}


class Program:
    """Procedures of the given reference source files (paths relative to REFERENCE_ROOT), ready to call()."""

    def __init__(self, files, root=None, constants_from=("src/TypeDefs.f90",), num_procs=1, proc_rank=0, ignore_calls=()):
        self.root = REFERENCE_ROOT if root is None else root
        self.ignore_calls = IGNORED_CALLS | {c.lower() for c in ignore_calls}   # subroutines off the path (logging, MPI, ...)
        self.files, self.procs, self.by_file = [], {}, {}
        self.globals = {"numprocs": num_procs, "procrank": proc_rank, "mpi_double_precision": 0, "mpi_comm_world": 0,
                        "mpi_integer": 0, "mpi_sum": 0}
        self.generics, self.units = {}, {}
        for rel in constants_from:
            _, decls, _ = parse_file(os.path.join(self.root, rel))
            env = {}
            for d in decls:
                self._declare(d[2], env, set(), module_level=True)
        self.globals.update(numprocs=num_procs, procrank=proc_rank)     # over TypeDefs.f90's initial values (1 and 0)
        for rel in files:
            procs, decls, generics = parse_file(os.path.join(self.root, rel))
            for d in decls:                         # initialised module variables (logInit = .false.) and parameters
                self._declare(d[2], {}, set(), module_level=True)
            self.by_file[rel] = procs
            self.files.append(rel)
            for name, p in procs.items():
                self.procs.setdefault(name, p)      # first file listed wins for names that several modules define
            for name, specifics in generics.items():
                self.generics.setdefault(name, []).extend((rel, sp) for sp in specifics)

    def where(self, name, file=None):
        p = self._find(name.lower(), file)
        return f"{os.path.relpath(p.file, self.root)}:{p.line}"

    def _decls_of(self, proc):
        if proc.decls is None:
            return [d for _, ln in proc.body if _DECL.match(ln) and "::" in ln for d in _parse_decl(ln)]
        return proc.decls

    def _find(self, name, file=None, nargs=None, first=None):
        """The procedure `name` as seen from `file`: a procedure of that file first, then a generic interface (the
        specific procedure with `nargs` dummies; generics of `file` first), then a procedure of any loaded file."""
        here = None
        if file is not None:
            for rel, procs in self.by_file.items():
                if rel == file or os.path.join(self.root, rel) == file:
                    here = rel
                    if name in procs:
                        return procs[name]
        if name in self.generics:
            cands = sorted(self.generics[name], key=lambda c: c[0] != here)
            fits = [p for p in (self.by_file[rel].get(sp) for rel, sp in cands)
                    if p is not None and (nargs is None or len(p.args) == nargs)]
            if len(fits) > 1 and isinstance(first, Obj) and first.tname:     # New(particles, ...) / New(faces, ...): by type
                for p in fits:
                    t = [d[4] for d in self._decls_of(p) if p.args and d[0] == p.args[0]]
                    if t and t[0] in (f"type({first.tname})", f"class({first.tname})"):
                        return p
            if fits:
                return fits[0]
        if name in self.procs:
            return self.procs[name]
        raise FortranError(f"procedure {name} not found in {self.files}")

    # -- declarations: constants and local arrays of constant size
    def _declare(self, ents, env, dummies, module_level=False, file=None):
        types = env.setdefault("%types", {}) if not module_level else {}
        for name, dims, init, is_param, base in ents:
            types[name] = base
            if (is_param or module_level) and init is not None:
                try:
                    val = self._eval(parse_expr(init), env, file)
                except (KeyError, FortranError, TypeError):
                    continue                        # kind(0.d0) and the like: not needed on the path
                (self.globals if module_level else env)[name] = val
            elif dims is not None and name not in dummies and not module_level:
                d = _split_top(dims)
                if all(":" not in x for x in d) and len(d) in (1, 2):
                    try:
                        n = [int(self._eval(parse_expr(x), env, file)) for x in d]
                    except (KeyError, FortranError):
                        continue
                    zero = _zero_of(base)
                    env[name] = FArr.zeros(n[0], value=zero) if len(n) == 1 else FMat(n[0], n[1], [zero] * (n[0] * n[1]))

    # -- expressions
    def _eval(self, e, env, file=None):
        k = e[0]
        if k == "lit":
            return e[1]
        if k == "paren":
            return self._eval(e[1], env, file)
        if k == "ref":
            return self._ref(e[1], env, file)
        if k == "array":
            return FArr([self._eval(x, env, file) for x in e[1]])
        if k == "neg":
            return _elementwise(lambda x: -x, self._eval(e[1], env, file))
        if k == "bin":
            a, b = self._eval(e[2], env, file), self._eval(e[3], env, file)
            op = e[1]
            if type(a) is float and type(b) is float:       # the common case, without the array dispatch
                if op == "+":
                    return a + b
                if op == "-":
                    return a - b
                if op == "*":
                    return a * b
                return _fdiv(a, b)
            if op == "+":
                return _elementwise(lambda x, y: x + y, a, b)
            if op == "-":
                return _elementwise(lambda x, y: x - y, a, b)
            if op == "*":
                return _elementwise(lambda x, y: x * y, a, b)
            return _elementwise(lambda x, y: _idiv(x, y) if isinstance(x, int) and isinstance(y, int)
                                and not isinstance(x, bool) else _fdiv(x, y), a, b)
        if k == "pow":
            a, b = self._eval(e[1], env, file), self._eval(e[2], env, file)

            def power(x, n):
                if isinstance(n, int) and not isinstance(n, bool):
                    if isinstance(x, int) and n >= 0:
                        return x ** n
                    # libgcc __powidf2: gfortran multiplies inline for n in -1 .. 2 and calls this for other integer
                    # powers (inline expansion of larger n needs -funsafe-math-optimizations); n = 2 gives x * x either way
                    m = abs(n)
                    y = x if m % 2 else 1.0
                    m >>= 1
                    while m:
                        x = x * x
                        if m % 2:
                            y = y * x
                        m >>= 1
                    return y if n >= 0 else _fdiv(1.0, y)
                try:
                    return math.pow(x, n)
                except (OverflowError, ValueError):
                    return float("nan")
            return _elementwise(power, a, b)
        if k == "rel":
            a, b = self._eval(e[2], env, file), self._eval(e[3], env, file)
            return {"==": a == b, "/=": a != b, "<": a < b, "<=": a <= b, ">": a > b, ">=": a >= b}[e[1]]
        if k == "and":
            return bool(self._eval(e[1], env, file)) and bool(self._eval(e[2], env, file))
        if k == "or":
            return bool(self._eval(e[1], env, file)) or bool(self._eval(e[2], env, file))
        if k == "not":
            return not self._eval(e[1], env, file)
        raise FortranError(f"cannot evaluate {e!r}")

    def _index(self, val, args, env, file):
        if isinstance(val, FMat):
            if len(args) != 2 or args[1][0] == "sec":
                raise FortranError("rank-2 arrays: a(i, j) and a(lo:hi, j) only")
            j = self._eval(args[1], env, file)
            if args[0][0] == "sec":
                return val.column(None if args[0][1] is None else self._eval(args[0][1], env, file),
                                  None if args[0][2] is None else self._eval(args[0][2], env, file), j)
            return val.get(self._eval(args[0], env, file), j)
        if callable(val):                           # a procedure dummy argument (topoFn)
            return val(*[self._eval(a, env, file) for a in args])
        if not isinstance(val, FArr):
            raise FortranError(f"subscript on a non-array value {val!r}")
        if len(args) != 1:
            raise FortranError("only rank-1 arrays are supported")
        a = args[0]
        if a[0] == "sec":
            return val.section(None if a[1] is None else self._eval(a[1], env, file),
                               None if a[2] is None else self._eval(a[2], env, file))
        return val.get(self._eval(a, env, file))

    def _ref(self, parts, env, file):
        name, args = parts[0]
        if name in env:
            val = env[name]
            if args is not None:
                val = self._index(val, args, env, file)
        elif name in self.globals and (args is None or isinstance(self.globals[name], (FArr, FMat))):
            val = self.globals[name]
            if args is not None:
                val = self._index(val, args, env, file)
        elif args is not None and name in ("allocated", "associated", "present"):
            try:                                    # a component that was never allocated does not exist here
                val = self._eval(args[0], env, file) is not None
            except (KeyError, FortranError):
                val = False
        elif args is not None and name in INTRINSICS:
            pos = [self._eval(a, env, file) for a in args if a[0] != "kw"]
            kw = {a[1][1][0][0]: self._eval(a[2], env, file) for a in args if a[0] == "kw"}
            val = INTRINSICS[name](*pos, **kw)
        elif args is not None:
            val = self._call(self._find(name, file, len(args)), args, env, file)
        else:
            try:
                proc = self._find(name, file)       # a procedure name passed as an actual argument
            except FortranError:
                raise KeyError(name)
            val = lambda *a, _p=proc: self._invoke(_p, list(a))
        for comp, cargs in parts[1:]:
            if not isinstance(val, Obj):
                raise FortranError(f"component {comp} of a non-derived-type value")
            try:
                val = val.f[comp]
            except KeyError:
                raise FortranError(f"derived type has no component {comp!r} (has {sorted(val.f)})")
            if cargs is not None:
                val = self._index(val, cargs, env, file)
        return val

    def _assign(self, lhs, value, env, file):
        if lhs[0] != "ref":
            raise FortranError("cannot assign to an expression")
        parts = lhs[1]
        # the container that holds the last part
        if len(parts) == 1:
            holder, key = env, parts[0][0]
        else:
            holder = self._ref(parts[:-1], env, file)
            if not isinstance(holder, Obj):
                raise FortranError("component assignment to a non-derived-type value")
            holder, key = holder.f, parts[-1][0]
        args = parts[-1][1]
        if args is not None:
            target = holder[key]
            if isinstance(target, FMat):
                if len(args) != 2 or args[1][0] == "sec":
                    raise FortranError("rank-2 arrays: a(i, j) and a(lo:hi, j) only")
                j = self._eval(args[1], env, file)
                if args[0][0] == "sec":
                    target.column(None if args[0][1] is None else self._eval(args[0][1], env, file),
                                  None if args[0][2] is None else self._eval(args[0][2], env, file), j).assign(value)
                else:
                    target.set(self._eval(args[0], env, file), j, value)
                return
            if len(args) != 1:
                raise FortranError("only rank-1 arrays are supported")
            a = args[0]
            if a[0] == "sec":
                target.section(None if a[1] is None else self._eval(a[1], env, file),
                               None if a[2] is None else self._eval(a[2], env, file)).assign(value)
            else:
                if isinstance(value, FArr):
                    raise FortranError("array assigned to an array element")
                target.set(self._eval(a, env, file), value)
            return
        cur = holder.get(key)
        if isinstance(cur, (FArr, FMat)):
            cur.assign(value)                       # whole-array assignment into existing storage
        elif isinstance(value, FArr):
            holder[key] = FArr(value.tolist())      # allocation on assignment (a copy)
        else:
            if isinstance(cur, float) and isinstance(value, int) and not isinstance(value, bool):
                value = float(value)                # a real variable keeps its type
            holder[key] = value

    # -- statements
    def _exec(self, stmts, env, file):
        for st in stmts:
            k = st[0]
            if k == "assign":
                self._assign(st[2], self._eval(st[3], env, file), env, file)
            elif k == "do":
                lo, hi = self._eval(st[3], env, file), self._eval(st[4], env, file)
                step = 1 if st[5] is None else self._eval(st[5], env, file)
                i = lo
                while (i <= hi) if step > 0 else (i >= hi):
                    env[st[2]] = i
                    try:
                        self._exec(st[6], env, file)
                    except _Cycle:
                        pass
                    except _Exit:
                        break
                    i += step
                else:
                    env[st[2]] = i
            elif k == "while":
                while self._eval(st[2], env, file):
                    self._exec(st[3], env, file)
            elif k == "if":
                for cond, body in st[2]:
                    if self._eval(cond, env, file):
                        self._exec(body, env, file)
                        break
                else:
                    if st[3] is not None:
                        self._exec(st[3], env, file)
            elif k == "call":
                if st[2] in self.ignore_calls:
                    continue
                first = None
                if st[3] and st[3][0][0] == "ref":
                    try:
                        first = self._eval(st[3][0], env, file)
                    except (KeyError, FortranError):
                        first = None
                self._call(self._find(st[2], file, len(st[3]), first), st[3], env, file)
            elif k == "allocate":               # allocate(a(n)): a fresh array of n zeros, lower bound 1
                parts = st[2][1]
                name, args = parts[-1]
                if args is None or len(args) not in (1, 2):
                    raise FortranError(f"line {st[1]}: unsupported allocate")
                holder = env if len(parts) == 1 else self._ref(parts[:-1], env, file).f
                zero = _zero_of(env.get("%types", {}).get(name, "real")) if len(parts) == 1 else 0.0
                if len(args) == 2:                  # allocate(a(n1, n2)); components are initialised by the caller's next statement
                    n1, n2 = (int(self._eval(a, env, file)) for a in args)
                    holder[name] = FMat(n1, n2, [zero] * (n1 * n2))
                elif args[0][0] == "sec":           # allocate(a(lo:hi))
                    lo, hi = int(self._eval(args[0][1], env, file)), int(self._eval(args[0][2], env, file))
                    holder[name] = FArr.zeros(hi - lo + 1, lb=lo, value=zero)
                else:
                    holder[name] = FArr.zeros(int(self._eval(args[0], env, file)), value=zero)
            elif k == "select":
                v = self._eval(st[2], env, file)
                for vals, body in st[3]:
                    if any(self._eval(x, env, file) == v for x in vals):
                        self._exec(body, env, file)
                        break
                else:
                    if st[4] is not None:
                        self._exec(st[4], env, file)
            elif k in ("open", "close", "read"):
                self._io(st, env, file)
            elif k == "return":
                raise _Return()
            elif k == "cycle":
                raise _Cycle()
            elif k == "exit":
                raise _Exit()
            elif k in ("decl", "continue"):
                continue
            else:
                raise FortranError(f"line {st[1]}: unsupported statement {st!r}")

    # -- list-directed input from the reference's own data files (the mesh seeds)
    def _io(self, st, env, file):
        kind, ctl, items = st[0], st[2], st[3]
        kw = {a[1][1][0][0]: a[2] for a in ctl if a[0] == "kw"}
        pos = [a for a in ctl if a[0] != "kw"]
        unit = self._eval(kw["unit"] if "unit" in kw else pos[0], env, file)
        if kind == "open":
            name = str(self._eval(kw["file"], env, file)).strip()
            path = name if os.path.isabs(name) else os.path.join(self.root, name)
            ok = os.path.isfile(path)
            if ok:
                recs = [re.split(r"[\s,]+", ln.strip()) for ln in open(path).read().splitlines()]
                self.units[unit] = [r for r in recs if r != [""]]
            if "iostat" in kw:
                self._assign(kw["iostat"], 0 if ok else 2, env, file)
            elif not ok:
                raise FortranError(f"cannot open {path}")
        elif kind == "close":
            self.units.pop(unit, None)
        else:                                       # read(unit, *) a, b, ...: a new record per statement, more as needed
            recs, toks = self.units[unit], []
            while len(toks) < len(items):
                if not recs:
                    raise FortranError("end of file")
                toks += recs.pop(0)
            for item, tok in zip(items, toks):
                try:
                    cur = self._eval(item, env, file)
                except (KeyError, FortranError):
                    cur = None
                if isinstance(cur, str) or (cur is None and env.get("%types", {}).get(item[1][0][0], "").startswith("character")):
                    val = tok
                elif isinstance(cur, float) or (cur is None and env.get("%types", {}).get(item[1][0][0], "real").startswith("real")):
                    val = float(tok.lower().replace("d", "e"))
                else:
                    val = int(tok)
                self._assign(item, val, env, file)

    def _call(self, proc, args, env, file):
        if len(args) > len(proc.args):
            raise FortranError(f"{proc.name}: {len(args)} arguments for {len(proc.args)} dummies")
        local, writeback = {}, []
        for dummy, a in zip(proc.args, args):
            val = self._eval(a, env, file)
            local[dummy] = val
            if not isinstance(val, (FArr, FMat, Obj)) and not callable(val) and a[0] == "ref" and (a[1][-1][1] is None or a[1][-1][1][0][0] != "sec"):
                is_var = a[1][0][0] in env or len(a[1]) > 1
                if is_var:
                    writeback.append((dummy, a))
        self._run(proc, local)
        for dummy, a in writeback:
            if dummy in local and not isinstance(local[dummy], (FArr, FMat, Obj)) and not callable(local[dummy]):
                self._assign(a, local[dummy], env, file)
        if proc.kind == "function":
            return local[proc.result]
        return None

    def _run(self, proc, local):
        if proc.decls is None:                      # first use: parse the body
            try:
                body, _, _ = parse_block(proc.body, 0, {"end", "endsubroutine", "endfunction"})
            except FortranError as e:
                raise FortranError(f"{os.path.relpath(proc.file, self.root)}:{proc.line} {proc.name}: {e}")
            proc.body = body
            proc.decls = [d for st in body if st[0] == "decl" for d in st[2]]
        self._declare(proc.decls, local, set(proc.args), file=proc.file)
        try:
            self._exec(proc.body, local, proc.file)
        except _Return:
            pass

    def _invoke(self, proc, values):
        local = dict(zip(proc.args, values))
        self._run(proc, local)
        return local[proc.result] if proc.kind == "function" else None

    def call(self, name, *actuals, file=None):
        """Calls procedure `name` with Python values (FArr / Obj / int / float / bool); returns the function result or,
        for a subroutine, the dict of dummy values after the call (scalars with intent(out) are read from it)."""
        proc = self._find(name.lower(), file)
        if len(actuals) != len(proc.args):
            raise FortranError(f"{proc.name} takes {len(proc.args)} arguments ({proc.args})")
        local = dict(zip(proc.args, actuals))
        self._run(proc, local)
        return local[proc.result] if proc.kind == "function" else local
