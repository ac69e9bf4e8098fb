"""A small interpreter for the Modern-Fortran subset the reference's hot path is written in.

TEST INFRASTRUCTURE (like everything under oracle/): it exists so that golden vectors for the gravity, sort-and-sweep
and drift routines can be produced by executing the REFERENCE'S OWN SOURCE TEXT, read from /root/reference/src at
generation time (tests/golden/gen_golden_fortran.py) -- no Fortran compiler exists in this image or on the GPU box
(profiles/r02_fortran_probe.txt).  Nothing of the reference is copied: the interpreter contains no Swiftest code, only
Fortran semantics.  The product never imports this module.

What it implements (enough for swiftest_kick.f90, swiftest_drift.f90, encounter_check.f90, encounter_util.f90
setup_aabb, the util_sort family of base_module.f90, swiftest_util_index_array, swiftest_orbel_scget):

* free-form source: continuation lines, comments, `#ifdef/#else/#endif` (macros undefined unless listed), OpenMP
  directives are comments (the loops run serially in statement order);
* modules / submodules: `parameter` constants, derived types (components, defaults, `extends`, type-bound procedures,
  `generic ::`), named generic interfaces (`module procedure`), procedures in `contains` parts; interface bodies are skipped;
* declarations: integer/real/logical/type/class with kinds, dimension (explicit, assumed shape, deferred), allocatable,
  optional, save, intent(out) (allocatable dummies are deallocated on entry, derived types re-initialised), automatic arrays;
* statements: assignment (array, section, vector subscript, (re)allocation on assignment), call (by reference: variables,
  array elements, contiguous or strided sections, components; keyword and optional arguments; generic and type-bound
  resolution by rank / type / kind), if, do, do while, do concurrent with mask, bare do, exit, cycle, return, where /
  elsewhere (masked evaluation of vector subscripts), associate, allocate (shape, source=, mold=), deallocate;
* expressions: Fortran precedence, integer division truncates, `x**n` with an integer n as repeated multiplication in
  the order gfortran expands it (x*x, (x*x)*x, binary method above), real `**` through libm pow, every operation an
  individually rounded IEEE double operation (Python floats / numpy float64 elementwise);
* intrinsics: size allocated present sqrt abs sum count pack merge any all int real min max mod sign sin cos norm2
  dot_product huge tiny epsilon move_alloc minval maxval.

`norm2` is processor dependent in the standard; NORM2_MODE selects "plain" (sqrt of the sum of squares, what compilers
inline and what -ffast-math gfortran -- the reference's release flags -- produces) or "libgfortran" (the scaled loop of
libgfortran's norm2_r8).  The golden generator runs both and records that the emitted pair lists are identical.
"""
import math
import re
import copy
import numpy as np

NORM2_MODE = "plain"


class FortranError(Exception):
    pass


# ------------------------------------------------------------------------------------------------ source handling
def preprocess(text, defines=()):
    out, stack = [], []
    for line in text.split("\n"):
        s = line.strip()
        if s.startswith("#"):
            m = re.match(r"#\s*(ifdef|ifndef|else|endif|if|elif|define|include|undef)\b\s*(\w*)", s)
            if not m:
                continue
            d, name = m.group(1), m.group(2)
            if d == "ifdef":
                stack.append(name in defines)
            elif d == "ifndef":
                stack.append(name not in defines)
            elif d == "if":
                stack.append(False)
            elif d == "else":
                stack[-1] = not stack[-1]
            elif d == "endif":
                stack.pop()
            continue
        if all(stack):
            out.append(line)
    return out


def strip_comment(line):
    q = None
    for k, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:k]
    return line


def logical_lines(text, defines=()):
    """-> list of (first physical line number, lowercased statement)."""
    res, cur, start = [], "", 0
    for no, raw in enumerate(preprocess(text, defines), 1):
        s = strip_comment(raw).strip()
        if not s:
            continue
        if cur:
            if s.startswith("&"):
                s = s[1:].lstrip()
        else:
            start = no
        if s.endswith("&"):
            cur += s[:-1].rstrip() + " "
            continue
        cur += s
        for part in split_semicolons(cur):
            res.append((start, lower_outside_strings(part.strip())))
        cur = ""
    return res


def split_semicolons(s):
    if ";" not in s:
        return [s]
    parts, q, cur = [], None, ""
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == ";":
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p for p in parts if p.strip()]


def lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
            out.append(ch.lower())
    return "".join(out)


# ------------------------------------------------------------------------------------------------ tokens / expressions
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<real>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)(?:_\w+)?|(?:\d+\.\d*|\.\d+)(?:_\w+)?)
  | (?P<int>\d+(?:_\w+)?)
  | (?P<dotop>\.(?:and|or|not|eqv|neqv|true|false|eq|ne|lt|le|gt|ge)\.(?:_\w+)?)
  | (?P<defop>\.[a-z]+\.)
  | (?P<name>[a-z_]\w*)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<op>\*\*|==|/=|<=|>=|=>|::|//|\(/|/\)|[-+*/<>=(),:%\[\]])
""", re.X)

DOT_REL = {".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m:
            raise FortranError("cannot tokenize: %r at %r" % (s, s[pos:pos + 20]))
        pos = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        v = m.group(k)
        if k == "dotop":
            v = v.split("_")[0] if v.startswith((".true.", ".false.")) else v
            if v in DOT_REL:
                toks.append(("op", DOT_REL[v]))
            elif v in (".true.", ".false."):
                toks.append(("log", v == ".true."))
            else:
                toks.append(("op", v))
        elif k == "defop":
            toks.append(("defop", v))
        elif k == "real":
            body = re.sub(r"_\w+$", "", v)
            toks.append(("num", float(body.replace("d", "e"))))
        elif k == "int":
            toks.append(("num", int(v.split("_")[0])))
        else:
            toks.append((k, v))
    return toks


class Parser:
    """Recursive descent over a token list; produces tuple ASTs."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        j = self.i + k
        return self.t[j] if j < len(self.t) else ("eof", None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, kind, val=None):
        tok = self.peek()
        if tok[0] == kind and (val is None or tok[1] == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        tok = self.next()
        if tok[0] != kind or (val is not None and tok[1] != val):
            raise FortranError("expected %s %r, got %r in %r" % (kind, val, tok, self.t))
        return tok

    def at_end(self):
        return self.i >= len(self.t)

    # precedence climbing, lowest first
    def expr(self):
        a = self.p_eqv()
        while self.peek()[0] == "defop":                 # defined binary operators bind loosest
            op = self.next()[1]
            a = ("defop", op, [a, self.p_eqv()])
        return a

    def p_eqv(self):
        a = self.p_or()
        while self.peek() in (("op", ".eqv."), ("op", ".neqv.")):
            op = self.next()[1]
            a = ("bin", op, a, self.p_or())
        return a

    def p_or(self):
        a = self.p_and()
        while self.peek() == ("op", ".or."):
            self.next()
            a = ("bin", ".or.", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek() == ("op", ".and."):
            self.next()
            a = ("bin", ".and.", a, self.p_not())
        return a

    def p_not(self):
        if self.peek() == ("op", ".not."):
            self.next()
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_cat()
        tok = self.peek()
        if tok[0] == "op" and tok[1] in ("==", "/=", "<", "<=", ">", ">="):
            self.next()
            a = ("bin", tok[1], a, self.p_cat())
        return a

    def p_cat(self):
        a = self.p_add()
        while self.peek() == ("op", "//"):
            self.next()
            a = ("bin", "//", a, self.p_add())
        return a

    def p_add(self):
        tok = self.peek()
        if tok in (("op", "-"), ("op", "+")):
            self.next()
            a = ("un", tok[1], self.p_mul())
        else:
            a = self.p_mul()
        while self.peek() in (("op", "-"), ("op", "+")):
            op = self.next()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.next()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek() == ("op", "**"):
            self.next()
            tok = self.peek()
            if tok in (("op", "-"), ("op", "+")):      # a ** -b
                self.next()
                b = ("un", tok[1], self.p_pow())
            else:
                b = self.p_pow()                        # right associative
            a = ("bin", "**", a, b)
        return a

    def p_primary(self):
        tok = self.next()
        k, v = tok
        if k == "num":
            return ("num", v)
        if k == "log":
            return ("num", v)
        if k == "str":
            return ("str", v[1:-1])
        if k == "op" and v == "(":
            e = self.expr()
            self.expect("op", ")")
            return ("paren", e)
        if k == "op" and v in ("[", "(/"):
            close = "]" if v == "[" else "/)"
            items = []
            if not self.accept("op", close):
                while True:
                    items.append(self.ac_item())
                    if self.accept("op", close):
                        break
                    self.expect("op", ",")
            return ("arr", items)
        if k == "name":
            node = ("name", v)
            return self.postfix(node)
        if k == "defop":                                  # defined unary operators bind tightest
            return ("defop", v, [self.p_primary()])
        raise FortranError("unexpected token %r in %r" % (tok, self.t))

    def ac_item(self):
        # implied do: ( expr-list , var = lo , hi [, step] )
        if self.peek() == ("op", "("):
            save = self.i
            try:
                self.next()
                exprs = [self.expr()]
                while self.accept("op", ","):
                    if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
                        var = self.next()[1]
                        self.next()
                        lo = self.expr()
                        self.expect("op", ",")
                        hi = self.expr()
                        st = self.expr() if self.accept("op", ",") else None
                        self.expect("op", ")")
                        return ("implied", exprs, var, lo, hi, st)
                    exprs.append(self.expr())
            except FortranError:
                pass
            self.i = save
        return self.expr()

    def postfix(self, node):
        while True:
            if self.peek() == ("op", "("):
                self.next()
                args = []
                if not self.accept("op", ")"):
                    while True:
                        args.append(self.arg())
                        if self.accept("op", ")"):
                            break
                        self.expect("op", ",")
                node = ("call", node, args)
            elif self.peek() == ("op", "%"):
                self.next()
                node = ("comp", node, self.expect("name")[1])
            else:
                return node

    def arg(self):
        if self.peek()[0] == "name" and self.peek(1) == ("op", "="):
            name = self.next()[1]
            self.next()
            return ("kw", name, self.expr())
        lo = None
        if self.peek() != ("op", ":"):
            lo = self.expr()
            if self.peek() != ("op", ":"):
                return lo
        self.expect("op", ":")
        hi = st = None
        if self.peek() not in (("op", ","), ("op", ")"), ("op", ":")):
            hi = self.expr()
        if self.accept("op", ":"):
            st = self.expr()
        return ("slice", lo, hi, st)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.at_end():
        raise FortranError("trailing tokens in expression %r" % s)
    return e


def split_top(s, sep=","):
    """Split at top-level separators (outside parentheses, brackets and strings)."""
    parts, depth, q, cur = [], 0, None, ""
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch in "([":
            depth += 1
            cur += ch
        elif ch in ")]":
            depth -= 1
            cur += ch
        elif depth == 0 and s.startswith(sep, i):
            parts.append(cur.strip())
            cur = ""
            i += len(sep)
            continue
        else:
            cur += ch
        i += 1
    parts.append(cur.strip())
    return parts


def match_paren(s, start):
    """s[start] == '(' -> index of the matching ')'."""
    depth, q = 0, None
    for i in range(start, len(s)):
        ch = s[i]
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
                return i
    raise FortranError("unbalanced parentheses in %r" % s)


# ------------------------------------------------------------------------------------------------ declarations
class Decl:
    __slots__ = ("base", "kind", "tname", "dims", "allocatable", "optional", "save", "intent", "parameter", "init", "name")

    def __init__(self):
        self.base = None; self.kind = None; self.tname = None; self.dims = None
        self.allocatable = self.optional = self.save = self.parameter = False
        self.intent = None; self.init = None; self.name = None

    def dtype(self):
        if self.base == "real":
            return np.float32 if self.kind in ("sp", "4") else np.float64
        if self.base == "integer":
            return np.int64 if self.kind in ("i8b", "8") else np.int32
        if self.base == "logical":
            return np.bool_
        return object

    @property
    def rank(self):
        return len(self.dims) if self.dims else 0


DECL_START = re.compile(r"^(integer|real|logical|character|complex|double\s+precision|type\s*\(|class\s*\()")


def parse_decl(stmt):
    """'real(dp), dimension(:), intent(in) :: a, b(3) = 0' -> [Decl]"""
    left, right = stmt.split("::", 1)
    left = left.strip()
    m = DECL_START.match(left)
    proto = Decl()
    word = m.group(1)
    rest = left[m.end():]
    if word.startswith(("type", "class")):
        close = match_paren(left, m.end() - 1)
        proto.base = "derived"
        proto.tname = left[m.end():close].strip()
        rest = left[close + 1:]
    else:
        proto.base = "real" if word.startswith("double") else word
        rest = rest.lstrip()
        if rest.startswith("("):
            close = match_paren(rest, 0)
            k = rest[1:close].strip()
            k = re.sub(r"^(kind|len)\s*=\s*", "", k)
            proto.kind = k
            rest = rest[close + 1:]
    for attr in split_top(rest):
        if not attr:
            continue
        if attr.startswith("dimension"):
            proto.dims = parse_dims(attr[attr.index("(") + 1:match_paren(attr, attr.index("("))])
        elif attr == "allocatable":
            proto.allocatable = True
        elif attr == "optional":
            proto.optional = True
        elif attr == "save":
            proto.save = True
        elif attr == "parameter":
            proto.parameter = True
        elif attr.startswith("intent"):
            proto.intent = re.sub(r"\s", "", attr[attr.index("(") + 1:-1])
        # target, pointer, contiguous, value, public, private: nothing to do here
    out = []
    for ent in split_top(right):
        d = copy.copy(proto)
        init = None
        if "=" in ent and "=>" not in ent:
            parts = split_top(ent, "=")
            if len(parts) == 2:
                ent, init = parts
        ent = ent.strip()
        m2 = re.match(r"^(\w+)\s*(\(.*\))?\s*(\*.*)?$", ent)
        if not m2:
            raise FortranError("cannot parse entity %r in %r" % (ent, stmt))
        d.name = m2.group(1)
        if m2.group(2):
            d.dims = parse_dims(m2.group(2)[1:-1])
        if init is not None:
            d.init = parse_expr(init)
            if not d.parameter:
                d.save = True
        out.append(d)
    return out


def parse_dims(s):
    dims = []
    for part in split_top(s):
        if part == ":":
            dims.append((None, None))
        elif part == "*":
            dims.append((None, "*"))
        else:
            lohi = split_top(part, ":")
            if len(lohi) == 1:
                dims.append((("num", 1), parse_expr(lohi[0])))
            else:
                dims.append((parse_expr(lohi[0]) if lohi[0] else ("num", 1), parse_expr(lohi[1]) if lohi[1] else None))
    return dims


# ------------------------------------------------------------------------------------------------ program structure
class Proc:
    __slots__ = ("name", "is_function", "args", "result", "lines", "decls", "body", "static", "file", "consts")

    def __init__(self):
        self.decls = None; self.body = None; self.static = {}; self.consts = None


class TypeDef:
    __slots__ = ("name", "parent", "comps", "bindings", "generics")

    def __init__(self, name, parent):
        self.name, self.parent = name, parent
        self.comps, self.bindings, self.generics = [], {}, {}


PROC_HEAD = re.compile(r"^(?:(?:pure|impure|elemental|recursive|module|non_recursive)\s+|"
                       r"(?:integer|real|logical|double\s+precision)\s*(?:\([^)]*\))?\s+)*"
                       r"(subroutine|function)\s+(\w+)\s*(\(.*?\))?\s*(?:result\s*\(\s*(\w+)\s*\))?\s*(?:bind\s*\(.*\))?$")
END_PROC = re.compile(r"^end\s*(subroutine|function)\b")
TYPE_HEAD = re.compile(r"^type\s*(?:,\s*(?P<attrs>[^:]*?))?\s*(?:::)?\s*(?P<name>\w+)$")


class World:
    """All the program units loaded so far: constants, types, generic interfaces, procedures."""

    def __init__(self, defines=()):
        self.defines = tuple(defines)
        self.consts = {}            # name -> value (evaluated lazily from const_exprs)
        self.const_exprs = {}       # name -> (Decl)
        self.types = {}
        self.generics = {}          # name -> [procedure names]
        self.procs = {}
        self.trace = None
        self.depth = 0

    # -- loading -------------------------------------------------------------------------------------------------
    def load(self, path):
        with open(path) as f:
            text = f.read()
        lines = logical_lines(text, self.defines)
        i, n = 0, len(lines)
        iface_depth = 0
        iface_name = None
        while i < n:
            no, s = lines[i]
            if re.match(r"^(abstract\s+)?interface\b", s):
                iface_depth += 1
                m = re.match(r"^interface\s+(\w+|operator\s*\(\s*\.[a-z]+\.\s*\))$", s)
                iface_name = re.sub(r"\s", "", m.group(1)) if m else None
                i += 1
                continue
            if re.match(r"^end\s*interface\b", s):
                iface_depth -= 1
                iface_name = None
                i += 1
                continue
            if iface_depth:
                m = re.match(r"^module\s+procedure\s*(?:::)?\s*(.*)$", s)
                if m and iface_name:
                    self.generics.setdefault(iface_name, []).extend(x.strip() for x in m.group(1).split(","))
                m = PROC_HEAD.match(s)
                if m and iface_name and not s.startswith("end"):      # interface body inside a named (generic) interface
                    self.generics.setdefault(iface_name, []).append(m.group(2))
                i += 1
                continue
            m = PROC_HEAD.match(s)
            if m:
                p = Proc()
                p.is_function = m.group(1) == "function"
                p.name = m.group(2)
                p.args = [a.strip() for a in m.group(3)[1:-1].split(",")] if m.group(3) and m.group(3)[1:-1].strip() else []
                p.result = m.group(4) or (p.name if p.is_function else None)
                p.file = path
                j = i + 1
                depth = 1
                while j < n:
                    sj = lines[j][1]
                    if PROC_HEAD.match(sj) and not sj.startswith("end"):
                        depth += 1
                    elif END_PROC.match(sj):
                        depth -= 1
                        if depth == 0:
                            break
                    j += 1
                p.lines = lines[i + 1:j]
                self.procs[p.name] = p
                i = j + 1
                continue
            m = TYPE_HEAD.match(s)
            if m and not s.startswith("type("):
                attrs = m.group("attrs") or ""
                pm = re.search(r"extends\s*\(\s*(\w+)\s*\)", attrs)
                td = TypeDef(m.group("name"), pm.group(1) if pm else None)
                i += 1
                in_contains = False
                while not re.match(r"^end\s*type\b", lines[i][1]):
                    sj = lines[i][1]
                    if sj == "contains":
                        in_contains = True
                    elif in_contains:
                        mb = re.match(r"^procedure\s*(?:\(\w+\))?\s*(?:,[^:]*)?::\s*(.*)$", sj)
                        mg = re.match(r"^generic\s*(?:,[^:]*)?::\s*(\w+)\s*=>\s*(.*)$", sj)
                        if mb:
                            for item in split_top(mb.group(1)):
                                if "=>" in item:
                                    a, b = [x.strip() for x in item.split("=>")]
                                else:
                                    a = b = item.strip()
                                td.bindings[a] = b
                        elif mg:
                            td.generics.setdefault(mg.group(1), []).extend(x.strip() for x in mg.group(2).split(","))
                    elif DECL_START.match(sj) and "::" in sj:
                        try:
                            td.comps.extend(parse_decl(sj))
                        except FortranError:
                            pass
                    i += 1
                self.types[td.name] = td
                i += 1
                continue
            if DECL_START.match(s) and "::" in s and "parameter" in s.split("::")[0]:
                try:
                    for d in parse_decl(s):
                        self.const_exprs[d.name] = d
                except FortranError:
                    pass
            i += 1

    def const(self, name):
        if name in self.consts:
            return self.consts[name]
        d = self.const_exprs.get(name)
        if d is None:
            raise KeyError(name)
        v = Frame(self, None).eval(d.init)
        v = convert_scalar(v, d) if not d.dims else v
        self.consts[name] = v
        return v

    # -- types ---------------------------------------------------------------------------------------------------
    def all_comps(self, tname):
        td = self.types.get(tname)
        if td is None:
            return []
        return (self.all_comps(td.parent) if td.parent else []) + td.comps

    def new_object(self, tname):
        obj = FObj(tname)
        for d in self.all_comps(tname):
            obj.c[d.name] = self.default_value(d, None)
            obj.decl[d.name] = d
        return obj

    def default_value(self, d, frame):
        if d.allocatable:
            return None
        if d.dims:
            if frame is None:
                frame = Frame(self, None)
            shape, lbs = [], []
            for lo, hi in d.dims:
                l = int(frame.eval(lo)); h = int(frame.eval(hi))
                shape.append(max(h - l + 1, 0)); lbs.append(l)
            if d.base == "derived":
                arr = FObjArray([self.new_object(d.tname) for _ in range(int(np.prod(shape)))])
                return arr
            a = np.zeros(shape, dtype=d.dtype(), order="F")
            if any(l != 1 for l in lbs):
                raise FortranError("lower bounds other than 1 are not supported (%s)" % d.name)
            if d.init is not None:
                a[...] = frame.eval(d.init)
            return a
        if d.base == "derived":
            return self.new_object(d.tname)
        if d.init is not None:
            return convert_scalar((frame or Frame(self, None)).eval(d.init), d)
        return {"real": 0.0, "integer": 0, "logical": False}.get(d.base, None)

    def type_chain(self, tname):
        chain = []
        while tname is not None:
            chain.append(tname)
            td = self.types.get(tname)
            tname = td.parent if td is not None else None
        return chain

    def find_binding(self, tname, name):
        """-> list of candidate procedure names for obj%name."""
        td = self.types.get(tname)
        while td is not None:
            if name in td.generics:
                return [self.find_binding(tname, g)[0] for g in td.generics[name]]
            if name in td.bindings:
                return [td.bindings[name]]
            td = self.types.get(td.parent) if td.parent else None
        raise FortranError("no binding %s in type %s" % (name, tname))

    # -- calling -------------------------------------------------------------------------------------------------
    def prepare(self, p):
        if p.body is not None:
            return
        decls, k = {}, 0
        lines = p.lines
        while k < len(lines):
            s = lines[k][1]
            if DECL_START.match(s) and "::" in s:
                for d in parse_decl(s):
                    decls[d.name] = d
            elif s.startswith("character") and find_assign(s) < 0:
                m = re.match(r"^character\s*(\(.*?\))?\s*(.*)$", s)          # old-style: no '::'
                for d in parse_decl("character%s :: %s" % (m.group(1) or "", m.group(2))):
                    decls[d.name] = d
            elif s.startswith(("implicit", "use ", "import", "external", "intrinsic")):
                pass
            else:
                break
            k += 1
        p.decls = decls
        p.body, k2 = parse_block(lines, k, ())
        if k2 != len(lines):
            raise FortranError("%s: could not parse %r (line %d)" % (p.name, lines[k2][1], lines[k2][0]))

    def resolve(self, cands, cells):
        """Pick the specific procedure whose dummies match the actual arguments (rank, type, integer array kind)."""
        if len(cands) == 1:
            return self.procs[cands[0]]
        for c in cands:
            p = self.procs.get(c)
            if p is None:
                continue
            self.prepare(p)
            if len(cells) > len(p.args):
                continue
            ok = True
            for a, (kw, cell) in zip(p.args, cells):
                d = p.decls[kw or a]
                v = cell.v if cell is not None else None
                if v is None:
                    ad = cell.decl if cell is not None else None      # unallocated actual: go by its declaration
                    if ad is not None and (ad.base != d.base or ad.rank != d.rank or
                                           (ad.base in ("integer", "real") and ad.dims and ad.dtype() != d.dtype())):
                        ok = False
                elif isinstance(v, np.ndarray):
                    if d.rank != v.ndim or d.base not in ("real", "integer", "logical") or np.dtype(d.dtype()) != v.dtype:
                        ok = False
                elif isinstance(v, (FObj, FObjArray)):
                    ok = ok and d.base == "derived"
                else:
                    if d.rank != 0:
                        ok = False
                    elif isinstance(v, bool):
                        ok = ok and d.base == "logical"
                    elif isinstance(v, float):
                        ok = ok and d.base == "real"
                    elif isinstance(v, int):
                        ok = ok and d.base == "integer"
                if not ok:
                    break
            for a in p.args[len(cells):]:
                if not p.decls[a].optional:
                    ok = False
            if ok:
                return p
        raise FortranError("no specific procedure among %s matches the arguments" % (cands,))

    def call(self, name, *values, **kw):
        """Python entry point: plain values / numpy arrays / Cells in, list of final argument values out."""
        cells = [(None, v if isinstance(v, CellBase) else Cell(v))for v in values]
        cells += [(k, v if isinstance(v, CellBase) else Cell(v)) for k, v in kw.items()]
        cands = self.generics.get(name, [name])
        p = self.resolve(cands, cells)
        res = self.invoke(p, cells)
        return res if p.is_function else [c.v for _, c in cells]

    def invoke(self, p, cells):
        self.prepare(p)
        fr = Frame(self, p)
        bound = set()
        pos = 0
        for kwname, cell in cells:
            if kwname is None:
                dummy = p.args[pos]; pos += 1
            else:
                dummy = kwname
            bound.add(dummy)
            d = p.decls.get(dummy)
            if cell is None:        # absent optional passed on
                continue
            if d is not None and d.intent == "out":
                if d.allocatable:
                    cell.v = None
                elif d.base == "derived" and isinstance(cell.v, FObj):
                    fresh = self.new_object(cell.v.tname)
                    cell.v.c, cell.v.decl = fresh.c, fresh.decl
            fr.vars[dummy] = cell
            if d is not None and getattr(cell, "decl", None) is None and isinstance(cell, Cell):
                cell.decl = d
        for name, d in p.decls.items():
            if name in bound or name in p.args:
                continue
            if d.parameter:
                fr.vars[name] = Cell(convert_scalar(fr.eval(d.init), d) if not d.dims else fr.eval(d.init), d)
                continue
            if d.save:
                if name not in p.static:
                    p.static[name] = Cell(self.default_value(d, fr), d)
                fr.vars[name] = p.static[name]
                continue
            fr.vars[name] = Cell(self.default_value(d, fr), d)
        if self.trace is not None:
            self.trace.append(("  " * self.depth) + p.name)
        self.depth += 1
        try:
            fr.run(p.body)
        except ReturnSignal:
            pass
        finally:
            self.depth -= 1
        if p.is_function:
            return fr.vars[p.result].v
        return None


# ------------------------------------------------------------------------------------------------ statements
def parse_block(lines, k, terminators):
    """Parse statements from lines[k:] until one starts with a terminator; -> (stmts, index of the terminator)."""
    out = []
    while k < len(lines):
        no, s = lines[k]
        if terminators and any(re.match(t, s) for t in terminators):
            return out, k
        stmt, k = parse_stmt(lines, k)
        if stmt is not None:
            out.append(stmt)
    return out, k


def strip_label(s):
    m = re.match(r"^(\w+)\s*:\s*(do|if|where|associate)\b(.*)$", s)
    if m and m.group(1) not in ("do", "if"):
        return m.group(2) + m.group(3)
    return s


def parse_stmt(lines, k):
    no, s = lines[k]
    s = strip_label(s)
    try:
        return _parse_stmt(lines, k, s)
    except FortranError as e:
        raise FortranError("line %d: %s  [%s]" % (no, s, e))


def _parse_stmt(lines, k, s):
    # ---- block constructs
    if re.match(r"^if\s*\(", s):
        close = match_paren(s, s.index("("))
        cond = parse_expr(s[s.index("(") + 1:close])
        tail = s[close + 1:].strip()
        if tail == "then":
            branches = []
            body, k = parse_block(lines, k + 1, (r"^else\b", r"^end\s*if\b", r"^elseif\b"))
            branches.append((cond, body))
            while True:
                t = lines[k][1]
                if re.match(r"^end\s*if\b", t):
                    return ("if", branches), k + 1
                m = re.match(r"^else\s*if\s*\(", t)
                if m:
                    o = t.index("(")
                    c2 = parse_expr(t[o + 1:match_paren(t, o)])
                    body, k = parse_block(lines, k + 1, (r"^else\b", r"^end\s*if\b", r"^elseif\b"))
                    branches.append((c2, body))
                else:
                    body, k = parse_block(lines, k + 1, (r"^end\s*if\b",))
                    branches.append((None, body))
        inner, _ = _parse_stmt([(0, tail)], 0, tail)
        return ("if", [(cond, [inner])]), k + 1
    if s == "do":
        body, k = parse_block(lines, k + 1, (r"^end\s*do\b",))
        return ("doforever", body), k + 1
    m = re.match(r"^do\s+while\s*\(", s)
    if m:
        o = s.index("(")
        cond = parse_expr(s[o + 1:match_paren(s, o)])
        body, k = parse_block(lines, k + 1, (r"^end\s*do\b",))
        return ("dowhile", cond, body), k + 1
    m = re.match(r"^do\s+concurrent\s*\(", s)
    if m:
        o = s.index("(")
        inside = s[o + 1:match_paren(s, o)]
        parts = split_top(inside)
        ctrls, mask = [], None
        for part in parts:
            mm = re.match(r"^(?:integer\s*(?:\(\w+\))?\s*::\s*)?(\w+)\s*=\s*(.*)$", part)
            if mm and ":" in part:
                rng = split_top(mm.group(2), ":")
                ctrls.append((mm.group(1), parse_expr(rng[0]), parse_expr(rng[1]), parse_expr(rng[2]) if len(rng) > 2 else None))
            else:
                mask = parse_expr(part)
        body, k = parse_block(lines, k + 1, (r"^end\s*do\b",))
        return ("doconc", ctrls, mask, body), k + 1
    m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s)
    if m:
        parts = split_top(m.group(2))
        body, k = parse_block(lines, k + 1, (r"^end\s*do\b",))
        return ("do", m.group(1), parse_expr(parts[0]), parse_expr(parts[1]),
                parse_expr(parts[2]) if len(parts) > 2 else None, body), k + 1
    if re.match(r"^where\s*\(", s):
        close = match_paren(s, s.index("("))
        mask = parse_expr(s[s.index("(") + 1:close])
        tail = s[close + 1:].strip()
        if tail:
            inner, _ = _parse_stmt([(0, tail)], 0, tail)
            return ("where", [(mask, [inner])]), k + 1
        branches = []
        body, k = parse_block(lines, k + 1, (r"^else\s*where\b", r"^end\s*where\b"))
        branches.append((mask, body))
        while True:
            t = lines[k][1]
            if re.match(r"^end\s*where\b", t):
                return ("where", branches), k + 1
            o = t.find("(")
            m2 = parse_expr(t[o + 1:match_paren(t, o)]) if o >= 0 else None
            body, k = parse_block(lines, k + 1, (r"^else\s*where\b", r"^end\s*where\b"))
            branches.append((m2, body))
    m = re.match(r"^select\s+type\s*\(", s)
    if m:
        o = s.index("(")
        inside = s[o + 1:match_paren(s, o)]
        if "=>" in inside:
            a, b = inside.split("=>", 1)
            name, sel = a.strip(), parse_expr(b.strip())
        else:
            name, sel = inside.strip(), parse_expr(inside.strip())
        guards = []
        k += 1
        term = (r"^class\s+is\b", r"^type\s+is\b", r"^class\s+default\b", r"^end\s*select\b")
        while not re.match(r"^end\s*select\b", lines[k][1]):
            t = lines[k][1]
            mg = re.match(r"^(class|type)\s+is\s*\(\s*(\w+)\s*\)", t)
            if mg:
                guard = (mg.group(1), mg.group(2))
            elif re.match(r"^class\s+default\b", t):
                guard = ("default", None)
            else:
                raise FortranError("unexpected statement in select type: %r" % t)
            body, k = parse_block(lines, k + 1, term)
            guards.append((guard, body))
        return ("selecttype", name, sel, guards), k + 1
    m = re.match(r"^select\s+case\s*\(", s)
    if m:
        o = s.index("(")
        sel = parse_expr(s[o + 1:match_paren(s, o)])
        cases = []
        k += 1
        term = (r"^case\b", r"^end\s*select\b")
        while not re.match(r"^end\s*select\b", lines[k][1]):
            t = lines[k][1]
            if re.match(r"^case\s+default\b", t):
                vals = None
            else:
                o = t.index("(")
                vals = [Parser(tokenize(x)).arg() for x in split_top(t[o + 1:match_paren(t, o)])]
            body, k = parse_block(lines, k + 1, term)
            cases.append((vals, body))
        return ("selectcase", sel, cases), k + 1
    if re.match(r"^associate\s*\(", s):
        o = s.index("(")
        pairs = []
        for part in split_top(s[o + 1:match_paren(s, o)]):
            a, b = part.split("=>", 1)
            pairs.append((a.strip(), parse_expr(b.strip())))
        body, k = parse_block(lines, k + 1, (r"^end\s*associate\b",))
        return ("associate", pairs, body), k + 1
    # ---- simple statements
    if s == "return":
        return ("return",), k + 1
    if re.match(r"^exit\b", s):
        return ("exit",), k + 1
    if re.match(r"^cycle\b", s):
        return ("cycle",), k + 1
    if s == "continue" or s.startswith(("write", "print", "format", "!$")):
        return None, k + 1
    if re.match(r"^call\s", s):
        e = parse_expr(s[4:].strip())
        if e[0] == "call":
            return ("call", e[1], e[2]), k + 1
        return ("call", e, []), k + 1
    m = re.match(r"^(allocate|deallocate)\s*\(", s)
    if m:
        o = s.index("(")
        items = [Parser(tokenize(x)).arg() for x in split_top(s[o + 1:match_paren(s, o)])]
        return (m.group(1), items), k + 1
    if re.match(r"^(stop|error\s+stop)\b", s):
        return ("stop", s), k + 1
    # assignment
    pos = find_assign(s)
    if pos < 0:
        raise FortranError("unsupported statement")
    return ("assign", parse_expr(s[:pos]), parse_expr(s[pos + 1:])), k + 1


def find_assign(s):
    depth, q = 0, None
    for i, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        elif ch == "=" and depth == 0:
            if s[i + 1:i + 2] in ("=", ">") or s[i - 1] in "<>/=":
                continue
            return i
    return -1


# ------------------------------------------------------------------------------------------------ run-time values
class CellBase:
    decl = None


class Cell(CellBase):
    __slots__ = ("v", "decl")

    def __init__(self, v=None, decl=None):
        self.v, self.decl = v, decl


class ElemCell(CellBase):
    __slots__ = ("arr", "idx")

    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    @property
    def v(self):
        return self.arr[self.idx].item()

    @v.setter
    def v(self, val):
        self.arr[self.idx] = val


class CompCell(CellBase):
    __slots__ = ("obj", "name")

    def __init__(self, obj, name):
        self.obj, self.name = obj, name

    @property
    def v(self):
        return self.obj.c[self.name]

    @v.setter
    def v(self, val):
        self.obj.c[self.name] = val

    @property
    def decl(self):
        return self.obj.decl.get(self.name)


class FObj:
    __slots__ = ("tname", "c", "decl")

    def __init__(self, tname, **comps):
        self.tname, self.c, self.decl = tname, dict(comps), {}

    def __deepcopy__(self, memo):
        o = FObj(self.tname)
        o.c = {k: copy.deepcopy(v, memo) for k, v in self.c.items()}
        o.decl = self.decl
        return o


class FObjArray:
    __slots__ = ("items",)

    def __init__(self, items):
        self.items = items


class ReturnSignal(Exception):
    pass


class ExitSignal(Exception):
    pass


class CycleSignal(Exception):
    pass


def convert_scalar(v, d):
    if d is None or isinstance(v, np.ndarray):
        return v
    if d.base == "integer" and not isinstance(v, (int, np.integer)):
        return int(v)           # truncation toward zero
    if d.base == "integer":
        return int(v)
    if d.base == "real":
        return float(v)
    if d.base == "logical":
        return bool(v)
    return v


def is_int(x):
    return isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_))


def int_pow(x, n):
    """x**n for an integer n the way gfortran expands it (powi: binary method; x*x, (x*x)*x for 2, 3)."""
    if n < 0:
        return 1.0 / int_pow(x, -n) if not is_int(x) else (0 if abs(x) > 1 else x ** n)
    if n == 0:
        return x * 0 + 1
    if n == 1:
        return x
    if n == 2:
        return x * x
    if n == 3:
        return (x * x) * x
    result, base, first = None, x, True
    while n:
        if n & 1:
            result = base if result is None else result * base
        n >>= 1
        if n:
            base = base * base
    return result


def trunc_div(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        q = np.abs(a) // np.abs(b)
        return (q * np.sign(a) * np.sign(b)).astype(np.result_type(a, b))
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def norm2(a):
    vals = [float(x) for x in np.asarray(a).ravel(order="F")]
    if NORM2_MODE == "plain":
        acc = 0.0
        for x in vals:
            acc = acc + x * x
        return math.sqrt(acc)
    scale, result = 1.0, 0.0           # libgfortran generated/norm2_r8.c
    for x in vals:
        if x != 0.0:
            ax = abs(x)
            if scale < ax:
                val = scale / ax
                result = 1.0 + result * val * val
                scale = ax
            else:
                val = ax / scale
                result += val * val
    return scale * math.sqrt(result)


def f_sign(a, b):
    return abs(a) if (b > 0 or (b == 0 and math.copysign(1.0, b) > 0)) else -abs(a)


def f_sum(a):
    a = np.asarray(a)
    if a.dtype.kind == "f":
        acc = 0.0
        for x in a.ravel(order="F"):
            acc = acc + float(x)
        return acc
    return int(a.sum())


def f_dot(a, b):
    acc = 0.0
    for x, y in zip(np.asarray(a).ravel(order="F"), np.asarray(b).ravel(order="F")):
        acc = acc + float(x) * float(y)
    return acc


def elementwise(fn):
    def g(x):
        if isinstance(x, np.ndarray):
            return np.array([fn(float(t)) for t in x.ravel(order="F")]).reshape(x.shape, order="F")
        return fn(x)
    return g


INTRINSICS = {
    "sqrt": lambda x: np.sqrt(x) if isinstance(x, np.ndarray) else (math.sqrt(x) if x >= 0 else math.nan),
    "abs": lambda x: np.abs(x) if isinstance(x, np.ndarray) else abs(x),
    "sin": elementwise(math.sin), "cos": elementwise(math.cos), "tan": elementwise(math.tan),
    "exp": elementwise(math.exp), "log": elementwise(math.log),
    "sinh": elementwise(math.sinh), "cosh": elementwise(math.cosh),
    "acos": elementwise(math.acos), "asin": elementwise(math.asin), "atan": elementwise(math.atan),
    "atan2": math.atan2,
    "sum": f_sum, "norm2": norm2, "dot_product": f_dot,
    "count": lambda m: int(np.count_nonzero(m)),
    "any": lambda m: bool(np.any(m)), "all": lambda m: bool(np.all(m)),
    "pack": lambda a, m: np.array(np.asarray(a)[np.asarray(m, dtype=bool)]),
    "merge": lambda t, f, m: np.where(m, t, f) if isinstance(m, np.ndarray) else (t if m else f),
    "min": lambda *a: min(a), "max": lambda *a: max(a),
    "minval": lambda a: np.asarray(a).min().item(), "maxval": lambda a: np.asarray(a).max().item(),
    "mod": lambda a, b: math.fmod(a, b) if isinstance(a, float) or isinstance(b, float) else int(math.fmod(a, b)),
    "sign": f_sign,
    "huge": lambda x: float(np.finfo(np.float64).max) if isinstance(x, float) else 2147483647,
    "tiny": lambda x: float(np.finfo(np.float64).tiny),
    "epsilon": lambda x: float(np.finfo(np.float64).eps),
    "nint": lambda x: int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5)),
    "floor": lambda x: int(math.floor(x)), "ceiling": lambda x: int(math.ceil(x)),
    "dble": float, "float": float,
}


class Frame:
    def __init__(self, world, proc):
        self.w, self.p = world, proc
        self.vars = {}
        self.where_mask = None

    # ---- lookup ---------------------------------------------------------------------------------------------
    def lookup(self, name):
        c = self.vars.get(name)
        if c is not None:
            return c
        return None

    def value_of_name(self, name):
        c = self.vars.get(name)
        if c is not None:
            return c.v
        try:
            return self.w.const(name)
        except KeyError:
            raise FortranError("unknown name %r in %s" % (name, self.p.name if self.p else "<const>"))

    # ---- indices --------------------------------------------------------------------------------------------
    def index(self, args, arr):
        idx, fancy = [], False
        for dim, a in enumerate(args):
            if a[0] == "slice":
                lo = None if a[1] is None else int(self.eval(a[1])) - 1
                hi = None if a[2] is None else int(self.eval(a[2]))
                st = None if a[3] is None else int(self.eval(a[3]))
                if st is not None and st < 0:
                    lo = arr.shape[dim] - 1 if lo is None else lo
                    hi = None if hi is None or hi - 2 < 0 else hi - 2
                if hi is not None and lo is not None and st in (None, 1) and hi < lo:
                    hi = lo
                idx.append(slice(lo, hi, st))
            else:
                v = self.eval(a)
                if isinstance(v, np.ndarray):
                    fancy = True
                    idx.append(v.astype(np.int64) - 1)
                else:
                    iv = int(v) - 1
                    if iv < 0 or iv >= arr.shape[dim]:
                        raise FortranError("index %d out of bounds (dimension %d, extent %d)" % (iv + 1, dim + 1, arr.shape[dim]))
                    idx.append(iv)
        return tuple(idx), fancy

    def getitem(self, arr, args):
        idx, fancy = self.index(args, arr)
        if fancy and self.where_mask is not None:
            mask = self.where_mask
            pos = [k for k, ix in enumerate(idx) if isinstance(ix, np.ndarray)]
            if len(pos) == 1 and idx[pos[0]].shape == mask.shape:
                out = np.zeros(mask.shape, dtype=arr.dtype)
                sub = list(idx)
                sub[pos[0]] = idx[pos[0]][mask]
                out[mask] = arr[tuple(sub)]
                return out
        r = arr[idx]
        if isinstance(r, np.ndarray):
            return r
        return r.item()

    # ---- evaluation -----------------------------------------------------------------------------------------
    def eval(self, e):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "name":
            return self.value_of_name(e[1])
        if k == "paren":
            return self.eval(e[1])
        if k == "bin":
            return self.binop(e[1], e[2], e[3])
        if k == "un":
            v = self.eval(e[2])
            if e[1] == "-":
                return -v
            if e[1] == "+":
                return v
            return np.logical_not(v) if isinstance(v, np.ndarray) else (not v)
        if k == "call":
            return self.eval_call(e)
        if k == "comp":
            base = self.eval(e[1])
            if isinstance(base, FObjArray):
                return np.array([o.c[e[2]] for o in base.items])
            return base.c[e[2]]
        if k == "arr":
            return self.array_constructor(e[1])
        if k == "defop":
            cells = [(None, self.ref(a)) for a in e[2]]
            p = self.w.resolve(self.w.generics["operator(%s)" % e[1]], cells)
            return self.w.invoke(p, cells)
        if k == "str":
            return e[1]
        raise FortranError("cannot evaluate %r" % (e,))

    def array_constructor(self, items):
        vals = []
        for it in items:
            if it[0] == "implied":
                _, exprs, var, lo, hi, st = it
                lo, hi = int(self.eval(lo)), int(self.eval(hi))
                st = 1 if st is None else int(self.eval(st))
                saved = self.vars.get(var)
                cell = Cell(0)
                self.vars[var] = cell
                for x in range(lo, hi + (1 if st > 0 else -1), st):
                    cell.v = x
                    for ex in exprs:
                        vals.append(np.atleast_1d(self.eval(ex)))
                if saved is not None:
                    self.vars[var] = saved
                else:
                    del self.vars[var]
            else:
                vals.append(np.atleast_1d(np.asarray(self.eval(it))).ravel(order="F"))
        if not vals:
            return np.zeros(0)
        out = np.concatenate(vals)
        return out

    def binop(self, op, ea, eb):
        if op == ".and.":
            a = self.eval(ea)
            if not isinstance(a, np.ndarray) and not a:
                b = self.eval(eb)
                return np.logical_and(a, b) if isinstance(b, np.ndarray) else False
            b = self.eval(eb)
            return np.logical_and(a, b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else bool(a and b)
        if op == ".or.":
            a, b = self.eval(ea), self.eval(eb)
            return np.logical_or(a, b) if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else bool(a or b)
        a, b = self.eval(ea), self.eval(eb)
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            ai = is_int(a) or (isinstance(a, np.ndarray) and a.dtype.kind == "i")
            bi = is_int(b) or (isinstance(b, np.ndarray) and b.dtype.kind == "i")
            if ai and bi:
                return trunc_div(a, b)
            if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
                with np.errstate(all="ignore"):
                    return np.divide(a, b, dtype=np.float64)
            if b == 0:
                a = float(a)
                return math.nan if (a == 0 or a != a) else math.copysign(math.inf, a) * math.copysign(1.0, float(b))
            return a / b
        if op == "**":
            if is_int(b):
                return int_pow(a, int(b))
            if isinstance(a, np.ndarray):
                return np.power(a, b)
            return math.pow(a, b)
        if op == "==":
            return a == b
        if op == "/=":
            return a != b
        if op == "<":
            return a < b
        if op == "<=":
            return a <= b
        if op == ">":
            return a > b
        if op == ">=":
            return a >= b
        if op == ".eqv.":
            return a == b
        if op == ".neqv.":
            return a != b
        raise FortranError("operator %s" % op)

    def eval_call(self, e):
        base, args = e[1], e[2]
        if base[0] == "name":
            name = base[1]
            c = self.vars.get(name)
            if c is not None:
                v = c.v
                if isinstance(v, FObjArray):
                    return self.index_objarray(v, args)
                if v is None:
                    raise FortranError("%s is not allocated" % name)
                return self.getitem(v, args)
            return self.intrinsic_or_function(name, args)
        v = self.eval(base)
        if isinstance(v, FObjArray):
            return self.index_objarray(v, args)
        return self.getitem(v, args)

    def index_objarray(self, v, args):
        a = args[0]
        if a[0] == "slice":
            lo = 0 if a[1] is None else int(self.eval(a[1])) - 1
            hi = len(v.items) if a[2] is None else int(self.eval(a[2]))
            return FObjArray(v.items[lo:hi])
        return v.items[int(self.eval(a)) - 1]

    def intrinsic_or_function(self, name, args):
        if name == "present":
            c = self.vars.get(args[0][1])
            return c is not None
        if name == "allocated":
            return self.ref(args[0]).v is not None
        pos = [a for a in args if a[0] != "kw"]
        kws = {a[1]: a[2] for a in args if a[0] == "kw"}
        if name == "size":
            v = self.eval(pos[0])
            dim = kws.get("dim", pos[1] if len(pos) > 1 else None)
            if isinstance(v, FObjArray):
                return len(v.items)
            return int(v.size) if dim is None else int(v.shape[int(self.eval(dim)) - 1])
        if name == "int":
            v = self.eval(pos[0])
            if isinstance(v, np.ndarray):
                kind = kws.get("kind", pos[1] if len(pos) > 1 else None)
                dt = np.int64 if kind is not None and kind[0] == "name" and kind[1] == "i8b" else np.int32
                return np.trunc(v).astype(dt)
            return int(v)
        if name == "real":
            v = self.eval(pos[0])
            return v.astype(np.float64) if isinstance(v, np.ndarray) else float(v)
        if name == "sum" and (len(pos) == 2 or "mask" in kws) and "dim" not in kws:
            a = np.asarray(self.eval(pos[0]))
            m = np.asarray(self.eval(kws.get("mask", pos[1] if len(pos) > 1 else None)), dtype=bool)
            return f_sum(a[m] if m.ndim else (a if m else a[:0]))
        if name in ("sum", "count", "any", "all", "minval", "maxval") and (len(pos) > 1 or kws):
            raise FortranError("%s with dim/mask is not supported" % name)
        fn = INTRINSICS.get(name)
        if fn is not None and name not in self.w.procs:
            return fn(*[self.eval(a) for a in pos])
        if name in self.w.procs or name in self.w.generics:
            cells = [(a[1], self.ref(a[2])) if a[0] == "kw" else (None, self.ref(a)) for a in args]
            p = self.w.resolve(self.w.generics.get(name, [name]), cells)
            return self.w.invoke(p, cells)
        raise FortranError("unknown function or array %r" % name)

    # ---- references (for assignment targets and actual arguments) ---------------------------------------------
    def ref(self, e):
        k = e[0]
        if k == "name":
            c = self.vars.get(e[1])
            if c is not None:
                return c
            if self.p is not None and e[1] in self.p.args:
                return None                         # absent optional dummy passed on
            return Cell(self.value_of_name(e[1]))
        if k == "comp":
            base = self.eval(e[1])
            if isinstance(base, FObj):
                return CompCell(base, e[2])
            return Cell(self.eval(e))
        if k == "call":
            base, args = e[1], e[2]
            holder = None
            if base[0] == "name":
                c = self.vars.get(base[1])
                if c is None:
                    return Cell(self.eval(e))       # function result
                holder = c.v
            elif base[0] == "comp":
                holder = self.eval(base)
            else:
                return Cell(self.eval(e))
            if isinstance(holder, FObjArray):
                return Cell(self.index_objarray(holder, args))
            if holder is None:
                raise FortranError("reference to an unallocated array in %s" % (self.p.name if self.p else "?"))
            idx, fancy = self.index(args, holder)
            if fancy:
                return Cell(self.getitem(holder, args))
            if all(isinstance(ix, int) for ix in idx):
                return ElemCell(holder, idx)
            return Cell(holder[idx])                # a view: writes reach the parent array
        return Cell(self.eval(e))

    # ---- execution ----------------------------------------------------------------------------------------------
    def run(self, stmts):
        for st in stmts:
            self.exec(st)

    def exec(self, st):
        k = st[0]
        if k == "assign":
            self.assign(st[1], st[2])
        elif k == "if":
            for cond, body in st[1]:
                if cond is None or self.eval(cond):
                    self.run(body)
                    break
        elif k == "call":
            self.exec_call(st[1], st[2])
        elif k == "do":
            _, var, lo, hi, step, body = st
            lo, hi = int(self.eval(lo)), int(self.eval(hi))
            step = 1 if step is None else int(self.eval(step))
            cell = self.vars[var]
            n_iter = max((hi - lo + step) // step, 0)
            x = lo
            try:
                for _ in range(n_iter):
                    cell.v = x
                    try:
                        self.run(body)
                    except CycleSignal:
                        pass
                    x += step
                    cell.v = x
                else:
                    cell.v = x
            except ExitSignal:
                pass
        elif k == "doconc":
            _, ctrls, mask, body = st
            self.do_concurrent(ctrls, 0, mask, body)
        elif k == "dowhile":
            try:
                while self.eval(st[1]):
                    try:
                        self.run(st[2])
                    except CycleSignal:
                        pass
            except ExitSignal:
                pass
        elif k == "doforever":
            try:
                while True:
                    try:
                        self.run(st[1])
                    except CycleSignal:
                        pass
            except ExitSignal:
                pass
        elif k == "return":
            raise ReturnSignal()
        elif k == "exit":
            raise ExitSignal()
        elif k == "cycle":
            raise CycleSignal()
        elif k == "where":
            self.exec_where(st[1])
        elif k == "associate":
            saved = {}
            for name, ex in st[1]:
                saved[name] = self.vars.get(name)
                r = self.ref(ex)
                self.vars[name] = r if r is not None else Cell(None)
            try:
                self.run(st[2])
            finally:
                for name, old in saved.items():
                    if old is None:
                        del self.vars[name]
                    else:
                        self.vars[name] = old
        elif k == "selecttype":
            _, name, sel, guards = st
            r = self.ref(sel)
            obj = r.v
            chain = self.w.type_chain(obj.tname) if isinstance(obj, FObj) else []
            chosen = None
            for (kind, tname), body in guards:             # type is > class is (most derived) > class default
                if kind == "type" and chain and chain[0] == tname:
                    chosen = body
                    break
            if chosen is None:
                best = None
                for (kind, tname), body in guards:
                    if kind == "class" and tname in chain and (best is None or chain.index(tname) < best[0]):
                        best = (chain.index(tname), body)
                chosen = best[1] if best else next((b for (kd, _), b in guards if kd == "default"), None)
            if chosen is not None:
                saved = self.vars.get(name)
                self.vars[name] = r
                try:
                    self.run(chosen)
                finally:
                    if saved is None:
                        del self.vars[name]
                    else:
                        self.vars[name] = saved
        elif k == "selectcase":
            v = self.eval(st[1])
            default = None
            for vals, body in st[2]:
                if vals is None:
                    default = body
                    continue
                hit = False
                for a in vals:
                    if a[0] == "slice":
                        lo = None if a[1] is None else self.eval(a[1])
                        hi = None if a[2] is None else self.eval(a[2])
                        hit = hit or ((lo is None or v >= lo) and (hi is None or v <= hi))
                    else:
                        hit = hit or v == self.eval(a)
                if hit:
                    self.run(body)
                    break
            else:
                if default is not None:
                    self.run(default)
        elif k == "allocate":
            self.exec_allocate(st[1])
        elif k == "deallocate":
            for it in st[1]:
                if it[0] != "kw":
                    self.ref(it).v = None
        elif k == "stop":
            raise FortranError("STOP reached: %s" % st[1])
        else:
            raise FortranError("cannot execute %r" % (st,))

    def do_concurrent(self, ctrls, level, mask, body):
        var, lo, hi, step = ctrls[level]
        lo, hi = int(self.eval(lo)), int(self.eval(hi))
        step = 1 if step is None else int(self.eval(step))
        cell = self.vars.get(var)
        if cell is None:
            cell = self.vars[var] = Cell(0)
        for x in range(lo, hi + (1 if step > 0 else -1), step):
            cell.v = x
            if level + 1 < len(ctrls):
                self.do_concurrent(ctrls, level + 1, mask, body)
            elif mask is None or self.eval(mask):
                try:
                    self.run(body)
                except CycleSignal:
                    pass

    def exec_call(self, callee, args):
        cells = []
        if callee[0] == "name" and callee[1] == "move_alloc":
            src, dst = self.ref(args[0]), self.ref(args[1])
            dst.v = src.v
            src.v = None
            return
        if callee[0] == "name" and callee[1].startswith("ieee_") and callee[1] not in self.w.procs:
            return                                      # floating-point environment calls: nothing to do here
        for a in args:
            if a[0] == "kw":
                cells.append((a[1], self.ref(a[2])))
            else:
                cells.append((None, self.ref(a)))
        if callee[0] == "comp":
            obj = self.eval(callee[1])
            cands = self.w.find_binding(obj.tname, callee[2])
            cells.insert(0, (None, Cell(obj)))
        else:
            name = callee[1]
            cands = self.w.generics.get(name, [name])
            if cands == [name] and name not in self.w.procs:
                if name.startswith("ieee_"):
                    return                              # floating-point environment calls: nothing to do here
                raise FortranError("call to unknown procedure %s" % name)
        p = self.w.resolve(cands, cells)
        self.w.invoke(p, cells)

    def exec_allocate(self, items):
        source = mold = None
        for it in items:
            if it[0] == "kw":
                if it[1] == "source":
                    source = self.eval(it[2])
                elif it[1] == "mold":
                    mold = self.eval(it[2])
        for it in items:
            if it[0] == "kw":
                continue
            if it[0] == "call":
                target = self.ref(it[1])
                shape = [int(self.eval(a)) if a[0] != "slice" else int(self.eval(a[2])) - int(self.eval(a[1])) + 1 for a in it[2]]
                shape = [max(s, 0) for s in shape]
                d = target.decl
                if d is None:
                    raise FortranError("allocate: no declaration for %r" % (it[1],))
                if d.base == "derived":
                    target.v = FObjArray([self.w.new_object(d.tname) for _ in range(int(np.prod(shape)))])
                else:
                    target.v = np.zeros(shape, dtype=d.dtype(), order="F")
                    if source is not None:
                        target.v[...] = source
            else:
                target = self.ref(it)
                if source is not None:
                    target.v = copy.deepcopy(source) if not isinstance(source, np.ndarray) else np.array(source, order="F")
                elif mold is not None:
                    target.v = np.zeros_like(mold, order="F")
                else:
                    d = target.decl
                    target.v = self.w.new_object(d.tname) if d is not None and d.base == "derived" else None

    def assign(self, lhs, rhs_e, mask=None):
        rhs = self.eval(rhs_e)
        if lhs[0] == "name":
            cell = self.vars.get(lhs[1])
            if cell is None:
                raise FortranError("assignment to undeclared %s" % lhs[1])
            self.store_whole(cell, rhs, mask)
            return
        if lhs[0] == "comp":
            base = self.eval(lhs[1])
            if isinstance(base, FObjArray):
                for n, o in enumerate(base.items):
                    if mask is None or mask[n]:
                        o.c[lhs[2]] = convert_scalar(rhs[n] if isinstance(rhs, np.ndarray) else rhs, o.decl.get(lhs[2]))
                return
            self.store_whole(CompCell(base, lhs[2]), rhs, mask)
            return
        if lhs[0] == "call":
            base, args = lhs[1], lhs[2]
            holder = self.vars[base[1]].v if base[0] == "name" else self.eval(base)
            if holder is None:
                raise FortranError("assignment to a section of an unallocated array")
            idx, _ = self.index(args, holder)
            if mask is not None:
                view = holder[idx]
                view[mask] = rhs[mask] if isinstance(rhs, np.ndarray) else rhs
                return
            if isinstance(rhs, np.ndarray) and holder.dtype.kind == "i" and rhs.dtype.kind == "f":
                rhs = np.trunc(rhs)
            elif holder.dtype.kind == "i" and isinstance(rhs, float):
                rhs = int(rhs)
            holder[idx] = rhs
            return
        raise FortranError("cannot assign to %r" % (lhs,))

    def store_whole(self, cell, rhs, mask):
        cur = cell.v
        d = cell.decl
        if isinstance(cur, np.ndarray):
            if mask is not None:
                cur[mask] = rhs[mask] if isinstance(rhs, np.ndarray) else rhs
            elif isinstance(rhs, np.ndarray) and rhs.shape != cur.shape and d is not None and d.allocatable:
                cell.v = np.array(rhs, dtype=cur.dtype, order="F")
            else:
                cur[...] = rhs
        elif cur is None and d is not None and d.dims and isinstance(rhs, np.ndarray):
            cell.v = np.array(rhs, dtype=d.dtype(), order="F")          # allocation on assignment
        elif isinstance(rhs, (FObj, FObjArray)):
            cell.v = copy.deepcopy(rhs)
        else:
            if isinstance(rhs, np.ndarray):
                raise FortranError("array assigned to a scalar")
            cell.v = convert_scalar(rhs, d) if d is not None else rhs

    def exec_where(self, branches):
        done = None
        for mask_e, body in branches:
            if mask_e is not None:
                m = np.asarray(self.eval(mask_e), dtype=bool)
                eff = m if done is None else (m & ~done)
                done = m.copy() if done is None else (done | m)
            else:
                eff = ~done
            old = self.where_mask
            self.where_mask = eff
            try:
                for st in body:
                    if st[0] != "assign":
                        raise FortranError("only assignments are supported inside where")
                    self.assign(st[1], st[2], mask=eff)
            finally:
                self.where_mask = old


def load_world(src_root, files, defines=()):
    import os
    w = World(defines)
    for f in files:
        w.load(os.path.join(src_root, f))
    return w
