"""A translator from the Fortran 90 subset the reference RRTMG / McICA sources are written in to Python, so that the
reference's OWN SOURCE TEXT can be executed here, where no Fortran compiler exists (oracle/build_ref.sh exits 3).

Test infrastructure, like everything under oracle/: it produces the golden vectors of tests/golden/ that pin the C
restatement (oracle/*.c) by the reference's code itself rather than by a second hand transcription.  Nothing of the
reference is copied: the sources are read where they lie under /root/reference and translated in memory.

What is translated (and nothing more): modules with module variables and `contains`ed subroutines / functions, `use`
with `only` lists and renames, declarations (integer / real / logical / character, parameter, dimension, intent,
optional, allocatable, pointer, save), statement functions, assignments (scalar, element, whole array, section, array
constructor), do / if / else if / where-elsewhere constructs, one-line if, call with positional and keyword
arguments, allocate / deallocate, return / cycle / exit, error stop, write (dropped).  Semantics held to the
reference build with `real` promoted to 8 bytes: every real is an IEEE double (Python float / numpy float64),
expressions are evaluated in the order written, integer division truncates, default integers wrap at 32 bits
(the KISS generator relies on it), scalar dummies are copied in and out (the generated function returns them),
array dummies are views of the actual argument re-bounded to the dummy's declaration (sequence association when the
shapes differ and the actual is contiguous).
"""
import math
import re

import numpy as np

# ------------------------------------------------------------------------------------------------------------------
# run-time support of the generated code
# ------------------------------------------------------------------------------------------------------------------


class FA:
    """A Fortran array: numpy data in Fortran order plus the lower bound of every dimension."""
    __slots__ = ("a", "lb")

    def __init__(self, a, lb=None):
        self.a = a
        self.lb = tuple(lb) if lb is not None else (1,) * a.ndim

    @staticmethod
    def new(kind, dims):
        """dims: sequence of (lo, hi)."""
        shape = tuple(max(0, int(hi) - int(lo) + 1) for lo, hi in dims)
        dt = {"real": np.float64, "integer": np.int64, "logical": np.bool_}[kind]
        return FA(np.zeros(shape, dtype=dt, order="F"), tuple(int(lo) for lo, _ in dims))


def _i32(x):
    """Default-integer (32-bit) wrap-around."""
    x = int(x)
    return ((x + 0x80000000) & 0xFFFFFFFF) - 0x80000000


def _idiv(a, b):
    a, b = int(a), int(b)
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _ishft(k, n):
    k = int(k) & 0xFFFFFFFF
    n = int(n)
    r = (k << n) & 0xFFFFFFFF if n >= 0 else k >> (-n)
    return _i32(r)


def _mod(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        a, b = int(a), int(b)
        return a - _idiv(a, b) * b
    return math.fmod(a, b)


def _sign(a, b):
    return abs(a) if b >= 0 else -abs(a)


def _nint(x):
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def _data(x):
    return x.a if isinstance(x, FA) else x


def _section(lo, hi, step, lb):
    """Python slice of the Fortran section lo:hi:step of a dimension with lower bound lb (hi inclusive)."""
    step = int(step)
    if step > 0:
        return slice(None if lo is None else int(lo) - lb, None if hi is None else int(hi) - lb + 1, step)
    stop = None if hi is None else int(hi) - lb - 1
    return slice(None if lo is None else int(lo) - lb, stop if (stop is None or stop >= 0) else None, step)


def _bind(actual, kind, dims, name):
    """Array dummy: a view of the actual argument with the dummy's bounds.  dims: list of (lo, hi) with hi None for
    assumed shape / size."""
    if actual is None:
        return None
    a = _data(actual)
    if not isinstance(a, np.ndarray):
        raise TypeError(f"array dummy {name} got a scalar")
    if any(hi is None for _, hi in dims):   # assumed shape (or size): keep the actual's shape
        if len(dims) != a.ndim:
            if dims[-1][1] is None and all(hi is not None for _, hi in dims[:-1]):   # assumed size: x(n,*)
                lead = [int(hi) - int(lo) + 1 for lo, hi in dims[:-1]]
                a = np.reshape(a, lead + [-1], order="F")
            else:
                raise ValueError(f"rank mismatch binding {name}")
        return FA(a, [int(lo) if lo is not None else 1 for lo, _ in dims])
    shape = tuple(max(0, int(hi) - int(lo) + 1) for lo, hi in dims)
    if a.shape != shape:
        n = int(np.prod(shape))
        flat = a.reshape(-1, order="F") if a.flags.f_contiguous or a.ndim == 1 else None
        if flat is None or flat.size < n or (flat.size and not np.shares_memory(flat, a)):
            raise ValueError(f"cannot bind {name}: actual shape {a.shape}, dummy shape {shape}")
        a = flat[:n].reshape(shape, order="F")
    return FA(a, [int(lo) for lo, _ in dims])


def _bind_pointer(actual, name):
    """Pointer dummy: disassociated (None) or the actual itself, bounds included."""
    if actual is None or isinstance(actual, FA):
        return actual
    a = _data(actual)
    if not isinstance(a, np.ndarray):
        raise TypeError(f"pointer dummy {name} got a scalar")
    return FA(a)


class StopError(RuntimeError):
    pass


class NamedCycle(Exception):
    """`cycle name` of an outer, named do construct (the argument is the construct name)."""


class NamedExit(Exception):
    """`exit name` of an outer, named do construct."""


RUNTIME = {"np": np, "math": math, "FA": FA, "_i32": _i32, "_idiv": _idiv, "_ishft": _ishft, "_mod": _mod, "_sign": _sign,
           "_nint": _nint, "_data": _data, "_section": _section, "_bind": _bind, "_bind_pointer": _bind_pointer, "StopError": StopError, "NamedCycle": NamedCycle,
           "NamedExit": NamedExit}

# ------------------------------------------------------------------------------------------------------------------
# source -> statements
# ------------------------------------------------------------------------------------------------------------------


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


def _lower_outside_strings(s):
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


# the object-like macros of MAPL_Generic.h the SW driver uses, as oracle/ref_recipe/MAPL_Generic.h states them
# (_ASSERT / _FAIL / _RETURN / _VERIFY are handled as statements by the translator)
MAPL_OBJECT_MACROS = (("__RC__", "RC=STATUS); _VERIFY(STATUS"), ("_RC", "RC=STATUS); _VERIFY(STATUS"), ("_SUCCESS", "0"),
                      ("_FAILURE", "1"))


def _expand_macros(line, macros):
    """cpp function-like macros: NAME(actuals) -> body with the formals replaced as whole words."""
    if "_" in line:
        for name, body in MAPL_OBJECT_MACROS:
            line = re.sub(r"(?<![A-Za-z0-9_])%s(?![A-Za-z0-9_])" % name, body, line)
    for name, (formals, body) in macros.items():
        while True:
            m = re.search(r"\b%s\s*\(" % re.escape(name), line)
            if not m:
                break
            j = _match_paren(line, m.end() - 1)
            actuals = [a.strip() for a in split_top(line[m.end():j])]
            if len(actuals) != len(formals):
                raise SyntaxError(f"macro {name}: {len(actuals)} actuals for {len(formals)} formals")
            sub = dict(zip(formals, actuals))
            text = re.sub(r"\b(%s)\b" % "|".join(map(re.escape, formals)), lambda mo: sub[mo.group(1)], body)
            line = line[:m.start()] + text + line[j + 1:]
    return line


def statements(text, defines=()):
    """Preprocessed logical statements [(lineno, text)]: cpp conditionals evaluated, comments removed, continuation
    lines joined, case folded outside strings."""
    lines = text.splitlines()
    active, stack = True, []
    phys = []
    macros = {}
    for no, raw in enumerate(lines, 1):
        s = raw.strip()
        if s.startswith("#"):
            m = re.match(r"#\s*(ifdef|ifndef|else|endif|include|define|if|undef)\b\s*(\S*)", s)
            if not m:
                continue
            k, arg = m.group(1), m.group(2)
            if k in ("ifdef", "ifndef", "if"):
                stack.append(active)
                cond = (arg in defines) if k == "ifdef" else (arg not in defines) if k == "ifndef" else False
                active = active and cond
                stack.append(cond)
            elif k == "else":
                cond = stack.pop()
                outer = stack[-1]
                active = outer and not cond
                stack.append(not cond)
            elif k == "endif":
                stack.pop()
                active = stack.pop()
            elif k == "define" and active:
                mm = re.match(r"#\s*define\s+(\w+)\(([^)]*)\)\s+(.*)$", s)
                if mm:   # function-like macro (SW/src/rrtmg_sw_cldprmc.F90:6-7)
                    macros[mm.group(1)] = ([a.strip() for a in mm.group(2).split(",")], mm.group(3).strip())
            continue
        if active:
            phys.append((no, _expand_macros(_strip_comment(raw), macros)))
    out, cur, cur_no = [], "", None
    for no, l in phys:
        t = l.strip()
        if not t:
            continue
        if cur:
            if t.startswith("&"):
                t = t[1:].lstrip()
            cur += " " + t
        else:
            cur, cur_no = t, no
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        for part in _split_semicolons(cur):
            part = part.strip()
            if part:
                out.append((cur_no, _lower_outside_strings(part)))
        cur = ""
    return out


def _split_semicolons(s):
    parts, cur, q = [], [], None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return parts


def split_top(s, sep=","):
    """Split at `sep` outside parentheses, brackets and strings."""
    parts, cur, depth, q = [], [], 0, None
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch in "([":
            depth += 1
            cur.append(ch)
        elif ch in ")]":
            depth -= 1
            cur.append(ch)
        elif ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur).strip())
    return parts


# ------------------------------------------------------------------------------------------------------------------
# expressions
# ------------------------------------------------------------------------------------------------------------------
TOKEN = re.compile(r"""\s*(?:
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?) |
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*") |
    (?P<dotop>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.) |
    (?P<name>[a-z_]\w*) |
    (?P<op>\*\*|//|==|/=|<=|>=|=>|\(/|/\)|[-+*/<>=(),:%\[\]])
)""", re.X)


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m:
            if s[pos:].strip() == "":
                break
            raise SyntaxError(f"cannot tokenise {s[pos:pos + 30]!r} in {s!r}")
        pos = m.end()
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
    return toks


INTRINSIC_TYPES = {  # result type: 'arg' = type of the first argument
    "exp": "real", "log": "real", "alog": "real", "sqrt": "real", "abs": "arg", "min": "arg", "max": "arg", "amax1": "real",
    "amin1": "real", "int": "integer", "nint": "integer", "real": "real", "dble": "real", "float": "real", "mod": "arg",
    "sign": "arg", "sum": "arg", "any": "logical", "all": "logical", "count": "integer", "maxval": "arg", "minval": "arg",
    "size": "integer", "present": "logical", "iand": "integer", "ior": "integer", "ieor": "integer", "ishft": "integer",
    "not": "integer", "allocated": "logical", "associated": "logical", "floor": "integer", "ceiling": "integer",
    "trim": "character", "epsilon": "real", "tiny": "real", "huge": "arg", "log10": "real", "sin": "real", "cos": "real",
    "acos": "real", "asin": "real", "atan": "real", "tan": "real", "lbound": "integer", "ubound": "integer", "kind": "integer",
    "merge": "arg", "storage_size": "integer", "index": "integer", "len_trim": "integer", "adjustl": "character",
}


class Node:
    pass


class Num(Node):
    def __init__(self, text):
        t = re.sub(r"_\w+$", "", text)
        self.is_int = re.fullmatch(r"\d+", t) is not None
        self.text = t.replace("d", "e")

    def typ(self, sc):
        return "integer" if self.is_int else "real"

    def rank(self, sc):
        return 0

    def py(self, sc):
        return self.text if self.is_int else repr(float(self.text))


class Str(Node):
    def __init__(self, text):
        self.text = text

    def typ(self, sc):
        return "character"

    def rank(self, sc):
        return 0

    def py(self, sc):
        body = self.text[1:-1].replace(self.text[0] * 2, self.text[0])
        return repr(body)


class Logical(Node):
    def __init__(self, v):
        self.v = v

    def typ(self, sc):
        return "logical"

    def rank(self, sc):
        return 0

    def py(self, sc):
        return "True" if self.v else "False"


class Slice(Node):
    def __init__(self, lo, hi, step):
        self.lo, self.hi, self.step = lo, hi, step


class Kw(Node):
    def __init__(self, name, val):
        self.name, self.val = name, val


class Name(Node):
    def __init__(self, name):
        self.name = name

    def typ(self, sc):
        return sc.type_of(self.name)

    def rank(self, sc):
        return sc.rank_of(self.name)

    def py(self, sc):
        ref = sc.ref(self.name)
        if sc.rank_of(self.name) > 0:
            return ref + ".a"
        return ref


class Apply(Node):
    """name(args): array element / section, function call or intrinsic - decided by the scope."""

    def __init__(self, name, args):
        self.name, self.args = name, args

    def kind(self, sc):
        if sc.is_array(self.name):
            return "array"
        if sc.is_stmt_function(self.name):
            return "stmtfn"
        if sc.is_function(self.name):
            return "function"
        if self.name in INTRINSIC_TYPES:
            return "intrinsic"
        raise NameError(f"{sc.where()}: unknown entity {self.name}(...)")

    def typ(self, sc):
        k = self.kind(sc)
        if k == "array":
            return sc.type_of(self.name)
        if k == "stmtfn":
            return sc.type_of(self.name)
        if k == "function":
            return sc.function_type(self.name)
        t = INTRINSIC_TYPES[self.name]
        if t == "arg":
            ts = [a.typ(sc) for a in self.args if not isinstance(a, Kw)]
            return "real" if "real" in ts else ts[0]
        return t

    def rank(self, sc):
        k = self.kind(sc)
        if k == "array":
            return sum(1 for a in self.args if isinstance(a, Slice) or a.rank(sc) > 0)
        if k == "intrinsic":
            if self.name in ("sum", "any", "all", "count", "maxval", "minval", "size", "present", "allocated", "lbound", "ubound"):
                return 0
            return max([a.rank(sc) for a in self.args if not isinstance(a, (Kw, Slice))] or [0])
        return 0

    def index_py(self, sc):
        """x.a[...] for an array reference."""
        lbs = sc.lower_bounds(self.name)
        ref = sc.ref(self.name)
        parts = []
        for d, a in enumerate(self.args):
            lb = lbs[d] if d < len(lbs) else None
            lbtxt = str(lb) if isinstance(lb, int) else f"{ref}.lb[{d}]"

            def off(e):
                if lbtxt == "0":
                    return e
                return f"{e}-{lbtxt}" if re.fullmatch(r"\w+", e) else f"({e})-{lbtxt}"
            if isinstance(a, Slice):
                lo = off(a.lo.py(sc)) if a.lo is not None else ""
                hi = f"{off(a.hi.py(sc))}+1" if a.hi is not None else ""
                if a.step is not None:
                    # a strided section: with a negative stride the last element is hi (inclusive) and a stop of -1
                    # would wrap around, so the slice is built at run time (e.g. `X(:,LM:1:-1)`, SOL:6136)
                    lo_t = a.lo.py(sc) if a.lo is not None else "None"
                    hi_t = a.hi.py(sc) if a.hi is not None else "None"
                    parts.append(f"_section({lo_t}, {hi_t}, {a.step.py(sc)}, {lbtxt})")
                    continue
                parts.append(f"{lo}:{hi}")
            elif a.rank(sc) > 0:   # vector subscript
                parts.append(f"np.asarray({a.py(sc)})-{lbtxt}")
            else:
                parts.append(off(a.py(sc)))
        return f"{ref}.a[{', '.join(parts)}]"

    def py(self, sc):
        k = self.kind(sc)
        if k == "array":
            return self.index_py(sc)
        if k == "stmtfn":
            return f"{self.name}__sf({', '.join(a.py(sc) for a in self.args)})"
        if k == "function":
            return sc.call_function(self.name, self.args)
        return intrinsic_py(self.name, self.args, sc)


def intrinsic_py(name, args, sc):
    pos = [a for a in args if not isinstance(a, Kw)]
    kws = {a.name: a.val for a in args if isinstance(a, Kw)}
    p = [a.py(sc) for a in pos]
    arr = any(a.rank(sc) > 0 for a in pos)
    if name in ("exp", "log", "alog", "sqrt", "log10", "sin", "cos", "acos", "asin", "atan", "tan"):
        fn = {"alog": "log"}.get(name, name)
        return f"np.{fn}({p[0]})" if arr else f"math.{fn}({p[0]})"
    if name == "abs":
        return f"np.abs({p[0]})" if arr else f"abs({p[0]})"
    if name in ("min", "max", "amax1", "amin1"):
        fn = "min" if name in ("min", "amin1") else "max"
        if arr:
            out = p[0]
            for q in p[1:]:
                out = f"np.{'minimum' if fn == 'min' else 'maximum'}({out}, {q})"
            return out
        return f"{fn}({', '.join(p)})"
    if name == "int":
        return f"{p[0]}.astype(np.int64)" if arr else f"int({p[0]})"
    if name == "nint":
        return f"_nint({p[0]})"
    if name in ("floor", "ceiling"):
        return f"int(math.{'floor' if name == 'floor' else 'ceil'}({p[0]}))"
    if name in ("real", "dble", "float"):
        return f"np.asarray({p[0]}, dtype=np.float64)" if arr else f"float({p[0]})"
    if name == "mod":
        return f"_mod({p[0]}, {p[1]})"
    if name == "sign":
        return f"_sign({p[0]}, {p[1]})"
    if name in ("sum", "maxval", "minval"):
        fn = {"sum": "sum", "maxval": "max", "minval": "min"}[name]
        return f"np.{fn}({p[0]})"
    if name in ("any", "all"):
        return f"bool(np.{name}({p[0]}))"
    if name == "count":
        return f"int(np.count_nonzero({p[0]}))"
    if name == "size":
        base = pos[0]
        ref = sc.ref(base.name) if isinstance(base, Name) else None
        if len(p) > 1 or "dim" in kws:
            d = p[1] if len(p) > 1 else kws["dim"].py(sc)
            return f"{ref}.a.shape[{d}-1]"
        return f"{ref}.a.size"
    if name == "present":
        return f"({sc.ref(pos[0].name)} is not None)"
    if name in ("allocated", "associated"):
        return f"({sc.ref(pos[0].name)} is not None)"
    if name == "iand":
        return f"(int({p[0]}) & int({p[1]}))"
    if name == "ior":
        return f"(int({p[0]}) | int({p[1]}))"
    if name == "ieor":
        return f"_i32(int({p[0]}) ^ int({p[1]}))"
    if name == "ishft":
        return f"_ishft({p[0]}, {p[1]})"
    if name == "epsilon":
        return "2.220446049250313e-16"
    if name == "tiny":
        return "2.2250738585072014e-308"
    if name == "huge":
        return "2147483647" if pos[0].typ(sc) == "integer" else "1.7976931348623157e+308"
    if name in ("trim", "adjustl"):
        return f"{p[0]}.strip()"
    if name == "merge":
        return f"np.where({p[2]}, {p[0]}, {p[1]})" if arr else f"({p[0]} if {p[2]} else {p[1]})"
    if name == "lbound":
        return f"{sc.ref(pos[0].name)}.lb[{p[1]}-1]"
    if name == "ubound":
        r = sc.ref(pos[0].name)
        return f"({r}.lb[{p[1]}-1]+{r}.a.shape[{p[1]}-1]-1)"
    if name == "storage_size":
        return "64"
    raise NotImplementedError(f"intrinsic {name}")


class Un(Node):
    def __init__(self, op, x):
        self.op, self.x = op, x

    def typ(self, sc):
        return "logical" if self.op == "not" else self.x.typ(sc)

    def rank(self, sc):
        return self.x.rank(sc)

    def py(self, sc):
        if self.op == "not":
            return f"(np.logical_not({self.x.py(sc)}))" if self.x.rank(sc) else f"(not {self.x.py(sc)})"
        return f"({self.op}{self.x.py(sc)})"


class Bin(Node):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b

    def typ(self, sc):
        if self.op in ("==", "!=", "<", "<=", ">", ">=", "and", "or", "eqv", "neqv"):
            return "logical"
        if self.op == "//":
            return "character"
        ta, tb = self.a.typ(sc), self.b.typ(sc)
        return "real" if "real" in (ta, tb) else ta

    def rank(self, sc):
        return max(self.a.rank(sc), self.b.rank(sc))

    def py(self, sc):
        a, b = self.a.py(sc), self.b.py(sc)
        op = self.op
        if op == "/" and self.a.typ(sc) == "integer" and self.b.typ(sc) == "integer":
            if self.rank(sc):
                raise NotImplementedError("integer array division")
            return f"_idiv({a}, {b})"
        if op in ("and", "or"):
            if self.rank(sc):
                return f"np.logical_{op}({a}, {b})"
            return f"({a} {op} {b})"
        if op == "eqv":
            return f"(bool({a}) == bool({b}))"
        if op == "neqv":
            return f"(bool({a}) != bool({b}))"
        if op == "//":
            return f"(str({a}) + str({b}))"
        if op == "**" and self.b.typ(sc) == "integer" and self.a.typ(sc) == "real" and not self.rank(sc):
            return f"(float({a}) ** {b})"
        return f"({a} {op} {b})"


class Cons(Node):
    """(/ a, b, ... /) with optional implied-do items (expr, i = lo, hi)."""

    def __init__(self, items):
        self.items = items

    def typ(self, sc):
        return self.items[0].typ(sc) if not isinstance(self.items[0], tuple) else self.items[0][0].typ(sc)

    def rank(self, sc):
        return 1

    def py(self, sc):
        parts = []
        for it in self.items:
            if isinstance(it, tuple):
                e, var, lo, hi = it
                sc.push_temp_int(var)
                parts.append(f"*[{e.py(sc)} for {var} in range({lo.py(sc)}, {hi.py(sc)}+1)]")
                sc.pop_temp_int(var)
            elif it.rank(sc) > 0:
                parts.append(f"*np.ravel({it.py(sc)}, order='F')")
            else:
                parts.append(it.py(sc))
        dt = "np.int64" if self.typ(sc) == "integer" else "np.float64" if self.typ(sc) == "real" else "None"
        return f"np.array([{', '.join(parts)}], dtype={dt})"


class Parser:
    PREC = [("eqv", "neqv"), ("or",), ("and",), ("not",), ("==", "!=", "<", "<=", ">", ">="), ("//",), ("+", "-"), ("*", "/"),
            ("**",)]

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
            raise SyntaxError(f"expected {val!r}, found {self.peek()} in {self.t}")

    DOT = {".and.": "and", ".or.": "or", ".not.": "not", ".eqv.": "eqv", ".neqv.": "neqv", ".eq.": "==", ".ne.": "!=",
           ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}

    def op_at(self):
        k, v = self.peek()
        if k == "dotop":
            return self.DOT.get(v)
        if k == "op":
            return {"/=": "!="}.get(v, v)
        return None

    def expr(self, level=0):
        if level == len(self.PREC):
            return self.primary()
        ops = self.PREC[level]
        if ops == ("not",):
            if self.op_at() == "not":
                self.next()
                return Un("not", self.expr(level))
            return self.expr(level + 1)
        if ops == ("+", "-"):
            if self.op_at() in ("+", "-"):
                sign = self.next()[1]
                left = self.expr(level + 1)
                left = Un("-", left) if sign == "-" else left
            else:
                left = self.expr(level + 1)
        elif ops == ("**",):
            left = self.expr(level + 1)
            if self.op_at() == "**":
                self.next()
                # right associative; the exponent may carry a sign
                if self.op_at() in ("+", "-"):
                    sign = self.next()[1]
                    right = self.expr(level)
                    right = Un("-", right) if sign == "-" else right
                else:
                    right = self.expr(level)
                return Bin("**", left, right)
            return left
        else:
            left = self.expr(level + 1)
        while self.op_at() in ops:
            op = self.op_at()
            self.next()
            right = self.expr(level + 1)
            left = Bin(op, left, right)
        return left

    def primary(self):
        k, v = self.next()
        if k == "num":
            return Num(v)
        if k == "str":
            return Str(v)
        if k == "dotop" and v in (".true.", ".false."):
            return Logical(v == ".true.")
        if k == "op" and v == "(":
            e = self.expr()
            self.expect(")")
            return e
        if k == "op" and v in ("(/", "["):
            close = "/)" if v == "(/" else "]"
            items = []
            while True:
                items.append(self.cons_item())
                if self.accept(close):
                    break
                self.expect(",")
            return Cons(items)
        if k == "name":
            node = Name(v)
            if self.peek()[1] == "(":
                self.next()
                node = Apply(v, self.arglist())
            while self.peek()[1] == "%":
                raise NotImplementedError("derived-type components")
            return node
        raise SyntaxError(f"unexpected token {v!r} in {self.t}")

    def cons_item(self):
        if self.peek()[1] == "(":   # maybe an implied do
            save = self.i
            self.next()
            try:
                e = self.expr()
                if self.accept(","):
                    k, var = self.next()
                    if k == "name" and self.accept("="):
                        lo = self.expr()
                        self.expect(",")
                        hi = self.expr()
                        self.expect(")")
                        return (e, var, lo, hi)
            except SyntaxError:
                pass
            self.i = save
        return self.expr()

    def arglist(self):
        args = []
        if self.accept(")"):
            return args
        while True:
            args.append(self.arg())
            if self.accept(")"):
                return args
            self.expect(",")

    def arg(self):
        # keyword argument?
        if self.peek()[0] == "name" and self.i + 1 < len(self.t) and self.t[self.i + 1][1] == "=" and \
                (self.i + 2 >= len(self.t) or self.t[self.i + 2][1] != "="):
            name = self.next()[1]
            self.next()
            return Kw(name, self.expr())
        lo = None
        if self.peek()[1] != ":":
            lo = self.expr()
            if self.peek()[1] != ":":
                return lo
        self.expect(":")
        hi = step = None
        if self.peek()[1] not in (",", ")", ":"):
            hi = self.expr()
        if self.accept(":"):
            step = self.expr()
        return Slice(lo, hi, step)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.i != len(p.t):
        raise SyntaxError(f"trailing tokens in {s!r}: {p.t[p.i:]}")
    return e


# ------------------------------------------------------------------------------------------------------------------
# program structure
# ------------------------------------------------------------------------------------------------------------------
TYPE_RE = re.compile(r"^(integer|real|double precision|logical|character|type)\b")


class Var:
    def __init__(self, name, typ):
        self.name, self.typ = name, typ
        self.dims = None          # list of (lo_expr_text | None, hi_expr_text | None)
        self.param = None         # initial / parameter expression text
        self.intent = None
        self.optional = self.allocatable = self.pointer = self.save = self.is_param = False


def parse_dims(text):
    dims = []
    for d in split_top(text):
        d = d.strip()
        if d in (":", "*"):
            dims.append((None, None))
        elif ":" in split_top(d, ":")[0:1][0] or len(split_top(d, ":")) == 2:
            lo, hi = split_top(d, ":")
            dims.append((lo.strip() or None, None if hi.strip() in ("", "*") else hi.strip()))
        else:
            dims.append(("1", d))
    return dims


def parse_decl(stmt):
    """-> list of Var, or None when stmt is not a type declaration."""
    m = TYPE_RE.match(stmt)
    if not m:
        return None
    typ = m.group(1)
    rest = stmt[m.end():].lstrip()
    if typ == "type" and not rest.startswith("("):
        return None   # a derived-type definition, not a declaration
    if typ == "double precision":
        typ = "real"
    # kind / len selectors
    if rest.startswith("*"):
        rest = re.sub(r"^\*\s*(\d+|\(.*?\))", "", rest).lstrip()
    elif rest.startswith("("):
        depth, j = 0, 0
        for j, ch in enumerate(rest):
            depth += ch == "("
            depth -= ch == ")"
            if depth == 0:
                break
        sel = rest[:j + 1]
        rest = rest[j + 1:].lstrip()
        if typ == "type":
            typ = "type:" + sel[1:-1].strip()
    attrs = []
    if "::" in rest:
        head, ents = rest.split("::", 1)
        attrs = [a.strip() for a in split_top(head) if a.strip()]
    else:
        if rest.startswith(","):
            return None
        ents = rest
        # `real function f(x)` is a procedure heading, not a declaration
        if re.match(r"(function|subroutine)\b", ents):
            return None
    out = []
    dim_attr = None
    for a in attrs:
        if a.startswith("dimension"):
            dim_attr = a[a.index("(") + 1:a.rindex(")")]
    for ent in split_top(ents):
        if not ent:
            continue
        init = None
        if "=" in ent and "=>" not in ent:
            parts = split_top(ent, "=")
            if len(parts) >= 2:
                ent, init = parts[0].strip(), "=".join(parts[1:]).strip()
        elif "=>" in ent:
            ent = ent.split("=>")[0].strip()
        m2 = re.match(r"([a-z_]\w*)\s*(\((.*)\))?\s*(\*\s*\S+)?$", ent)
        if not m2:
            raise SyntaxError(f"cannot parse entity {ent!r} in {stmt!r}")
        v = Var(m2.group(1), typ)
        dtext = m2.group(3) if m2.group(2) else dim_attr
        if dtext is not None:
            v.dims = parse_dims(dtext)
        v.param = init
        for a in attrs:
            if a == "parameter":
                v.is_param = True
            elif a.startswith("intent"):
                v.intent = re.sub(r"\s", "", a)[7:-1]
            elif a == "optional":
                v.optional = True
            elif a == "allocatable":
                v.allocatable = True
            elif a == "pointer":
                v.pointer = True
            elif a == "save":
                v.save = True
        out.append(v)
    return out


class Proc:
    def __init__(self, name, kind, args, result, module, prefix_type=None):
        self.name, self.kind, self.args, self.result, self.module = name, kind, args, result, module
        self.vars = {}
        self.uses = []            # (module, {local: remote} or None)
        self.body = []            # [(lineno, stmt)]
        self.stmt_functions = {}  # name -> (args, expr text)
        self.prefix_type = prefix_type
        self.out_scalars = []     # scalar dummies copied back to the caller, in order


class Module:
    def __init__(self, name):
        self.name = name
        self.vars = {}
        self.uses = []
        self.procs = {}
        self.generics = {}


PROC_RE = re.compile(r"^(?:(?:pure|elemental|recursive)\s+)*(?:(integer|real|logical|double precision)\s+)?"
                     r"(subroutine|function)\s+([a-z_]\w*)\s*(\((.*?)\))?\s*(?:result\s*\(\s*([a-z_]\w*)\s*\))?$")


def parse_use(stmt):
    m = re.match(r"use\s*(?:,\s*intrinsic\s*)?(?:::)?\s*([a-z_]\w*)\s*(?:,\s*only\s*:\s*(.*))?$", stmt)
    if not m:
        raise SyntaxError(f"cannot parse {stmt!r}")
    only = None
    if m.group(2) is not None:
        only = {}
        for it in split_top(m.group(2)):
            if not it:
                continue
            if "=>" in it:
                loc, rem = [x.strip() for x in it.split("=>")]
            else:
                loc = rem = it.strip()
            only[loc] = rem
    return m.group(1), only


class Program:
    """All modules of a set of source files."""

    def __init__(self):
        self.modules = {}
        self.loose_procs = {}

    def add_source(self, text, fname="", defines=()):
        st = statements(text, defines)
        i = 0
        mod, in_iface, in_type = None, 0, False
        stack = []          # procedure nesting: [module procedure, internal procedure]
        while i < len(st):
            no, s = st[i]
            i += 1
            proc = stack[-1] if stack else None
            if in_iface:
                if re.match(r"end\s*interface", s):
                    in_iface -= 1
                continue
            if in_type:
                if re.match(r"end\s*type", s):
                    in_type = False
                continue
            if re.match(r"(abstract\s+)?interface\b", s):
                in_iface += 1
                continue
            if re.match(r"type\s*(,|::|\s+[a-z_])", s) and not s.startswith("type("):
                in_type = True
                continue
            m = re.match(r"module\s+([a-z_]\w*)$", s)
            if m and proc is None:
                mod = Module(m.group(1))
                mod.file = fname
                self.modules[mod.name] = mod
                continue
            if re.match(r"end\s*module", s):
                mod = None
                continue
            if re.match(r"end\s*(subroutine|function)", s) or (s == "end" and proc is not None):
                stack.pop()
                continue
            if s == "contains":
                if proc is not None:
                    proc.spec_done = True
                continue
            pm = PROC_RE.match(s)
            if pm and (proc is None or proc.spec_done):
                args = [a.strip() for a in split_top(pm.group(5))] if pm.group(5) else []
                args = [a for a in args if a]
                new = Proc(pm.group(3), pm.group(2), args, pm.group(6) or (pm.group(3) if pm.group(2) == "function" else None),
                           mod.name if mod else None, pm.group(1))
                new.file, new.line = fname, no
                new.spec_done = False
                new.host = proc
                new.internal = {}
                if proc is not None:
                    proc.internal[new.name] = new
                else:
                    (mod.procs if mod else self.loose_procs)[new.name] = new
                stack.append(new)
                continue
            owner = proc if proc is not None else mod
            if owner is None:
                continue
            if s.startswith("use ") or s.startswith("use,"):
                owner.uses.append(parse_use(s))
                continue
            if s.startswith("implicit") or re.match(r"(public|private|save|external|intrinsic)\b", s):
                continue
            if proc is None or not proc.spec_done:
                d = parse_decl(s)
                if d is not None:
                    for v in d:
                        if v.name in owner.vars:   # attributes given in separate statements: merge
                            old = owner.vars[v.name]
                            old.dims = old.dims or v.dims
                            old.typ = v.typ
                            old.param = old.param or v.param
                        else:
                            owner.vars[v.name] = v
                    continue
                m = re.match(r"parameter\s*\((.*)\)$", s)
                if m:
                    for it in split_top(m.group(1)):
                        n, e = it.split("=", 1)
                        owner.vars[n.strip()].param = e.strip()
                        owner.vars[n.strip()].is_param = True
                    continue
                if re.match(r"(dimension|intent\s*\(\s*\w+\s*\)|optional|allocatable|pointer|target)\s*(::|\s)", s):
                    continue
                if re.match(r"data\s", s):
                    owner.__dict__.setdefault("data", []).append((no, s))
                    continue
                if re.match(r"equivalence\b", s):   # (a(1,1,1), b(1,1)) [, (...)]: whole arrays sharing their storage
                    for grp in re.findall(r"\(\s*([a-z_]\w*)\s*\([\d,\s]*\)\s*,\s*([a-z_]\w*)\s*\([\d,\s]*\)\s*\)", s):
                        owner.__dict__.setdefault("equiv", []).append(grp)
                    continue
            if proc is not None:
                # statement function?  name(dummy, ...) = expr with name a declared SCALAR
                sf = re.match(r"([a-z_]\w*)\s*\(([a-z_\w\s,]*)\)\s*=(?!=)\s*(.*)$", s)
                if not proc.spec_done and sf and sf.group(1) in proc.vars and proc.vars[sf.group(1)].dims is None \
                        and sf.group(1) not in proc.args:
                    proc.stmt_functions[sf.group(1)] = ([a.strip() for a in sf.group(2).split(",")], sf.group(3))
                    continue
                proc.spec_done = True
                proc.body.append((no, s))
        return self


# ------------------------------------------------------------------------------------------------------------------
# code generation
# ------------------------------------------------------------------------------------------------------------------
PY_RESERVED = {"and", "as", "assert", "break", "class", "continue", "def", "del", "elif", "else", "except", "finally", "for",
               "from", "global", "if", "import", "in", "is", "lambda", "nonlocal", "not", "or", "pass", "raise", "return", "try",
               "while", "with", "yield", "none", "true", "false", "int", "float", "abs", "min", "max", "range", "str", "bool",
               "len", "np", "math", "sum", "any", "all", "round", "print", "slice", "tuple", "list", "fa", "type", "id", "map",
               "iter", "next", "set", "dict", "object", "input", "open", "exec", "eval"}


def pyname(n):
    return n + "_v" if n in PY_RESERVED else n


class Scope:
    """Name resolution inside one procedure (or a module's specification part when proc is None)."""

    def __init__(self, tr, module, proc):
        self.tr, self.module, self.proc = tr, module, proc
        self.temp_ints = []
        self.line = 0

    def where(self):
        return f"{self.module.name if self.module else '?'}::{self.proc.name if self.proc else '<module>'} line {self.line}"

    # ---- lookup: returns (kind, owner, Var|Proc) with kind in local / host / module / proc
    def lookup(self, name):
        p = self.proc
        while p is not None:
            if name in p.vars:
                return ("local" if p is self.proc else "host", p, p.vars[name])
            if name in p.internal:
                return ("proc", p, p.internal[name])
            p = p.host
        seen = set()
        # the procedure's own `use` statements, then its hosts', then the module's
        p = self.proc
        while p is not None:
            r = self._via_uses(p.uses, name, seen)
            if r:
                return r
            p = p.host
        if self.module is not None:
            r = self._in_module(self.module, name, seen)
            if r:
                return r
        if name in self.tr.program.loose_procs:   # external procedures (not in a module)
            return ("proc", None, self.tr.program.loose_procs[name])
        return None

    def _in_module(self, mod, name, seen):
        if mod.name in seen:
            return None
        seen.add(mod.name)
        if name in mod.vars:
            return ("module", mod, mod.vars[name])
        if name in mod.procs:
            return ("proc", mod, mod.procs[name])
        return self._via_uses(mod.uses, name, seen)

    def _via_uses(self, uses, name, seen):
        for mname, only in uses:
            m = self.tr.program.modules.get(mname)
            if m is None:
                continue
            if only is not None:
                if name in only:
                    r = self._in_module(m, only[name], set())
                    if r:
                        return r
                continue
            r = self._in_module(m, name, set(seen))
            if r:
                return r
        return None

    def push_temp_int(self, v):
        self.temp_ints.append(v)

    def pop_temp_int(self, v):
        self.temp_ints.remove(v)

    def var(self, name):
        r = self.lookup(name)
        if r is None or r[0] == "proc":
            return None
        return r[2]

    def type_of(self, name):
        if name in self.temp_ints:
            return "integer"
        if self.proc is not None and name in self.proc.stmt_functions:
            return self.proc.vars[name].typ
        v = self.var(name)
        if v is None:
            r = self.lookup(name)
            if r and r[0] == "proc":
                return self.function_type(name)
            raise NameError(f"{self.where()}: unknown name {name}")
        return "type" if v.typ.startswith("type") else v.typ

    def rank_of(self, name):
        if name in self.temp_ints:
            return 0
        v = self.var(name)
        if v is None:
            return 0
        return len(v.dims) if v.dims else 0

    def is_array(self, name):
        if self.proc is not None and name in self.proc.stmt_functions:
            return False
        v = self.var(name)
        return v is not None and v.dims is not None

    def is_stmt_function(self, name):
        p = self.proc
        while p is not None:
            if name in p.stmt_functions:
                return True
            p = p.host
        return False

    def is_function(self, name):
        r = self.lookup(name)
        return r is not None and r[0] == "proc" and r[2].kind == "function"

    def function_type(self, name):
        f = self.lookup(name)[2]
        if f.prefix_type:
            return "real" if f.prefix_type == "double precision" else f.prefix_type
        rv = f.vars.get(f.result)
        return rv.typ if rv else "real"

    def lower_bounds(self, name):
        v = self.var(name)
        out = []
        for lo, _ in v.dims:
            if lo is None:
                # deferred shape: a pointer carries the bounds of its target (known at run time only);
                # assumed-shape dummies and freshly allocated arrays without a lower bound start at 1
                out.append(None if v.pointer else 1)
            elif re.fullmatch(r"-?\d+", lo):
                out.append(int(lo))
            else:
                out.append(None)
        return out

    def ref(self, name):
        if name in self.temp_ints:
            return pyname(name)
        r = self.lookup(name)
        if r is None:
            raise NameError(f"{self.where()}: unknown name {name}")
        kind, owner, obj = r
        if kind in ("local", "host"):
            if getattr(obj, "static", False):
                return f"S_{owner.pyid}.{pyname(obj.name)}"
            return pyname(obj.name)
        if kind == "module":
            return f"M_{owner.name}.{pyname(obj.name)}"
        return self.tr.proc_pyname(obj)

    def call_function(self, name, args):
        f = self.lookup(name)[2]
        return self.tr.call_text(self, f, args, as_function=True)


class Translator:
    def __init__(self, program):
        self.program = program
        self.out = []
        for m in program.modules.values():
            for p in m.procs.values():
                self._prepare(p, m)
        for p in program.loose_procs.values():
            self._prepare(p, None)

    # ---- analysis ----
    def _prepare(self, p, mod, prefix=None):
        p.pyid = (prefix or (mod.name if mod else "x")) + "__" + p.name
        p.mod = mod
        for v in p.vars.values():
            # initialised or SAVEd locals are static (Fortran: initialisation implies SAVE)
            v.static = (v.name not in p.args) and not v.is_param and (v.param is not None or v.save) and v.name != p.result
        body_text = "\n".join(s for _, s in p.body)
        p.out_scalars = []
        for a in p.args:
            v = p.vars.get(a)
            if v is None or v.dims is not None:
                continue
            assigned = re.search(r"(^|\n|\)\s)\s*%s\s*=(?!=)" % re.escape(a), body_text) is not None
            passed = re.search(r"call\s+\w+\s*\(.*\b%s\b" % re.escape(a), body_text) is not None
            if v.intent in ("out", "inout") or (v.intent is None and (assigned or passed)):
                p.out_scalars.append(a)
        for q in p.internal.values():
            self._prepare(q, mod, p.pyid)

    def proc_pyname(self, p):
        return "P_" + p.pyid

    # ---- emit helpers ----
    def emit(self, ind, text):
        self.out.append("    " * ind + text)

    def dims_py(self, sc, v):
        parts = []
        for lo, hi in v.dims:
            lo_t = parse_expr(lo).py(sc) if lo is not None else "1"
            hi_t = parse_expr(hi).py(sc) if hi is not None else "None"
            parts.append(f"({lo_t}, {hi_t})")
        return "[" + ", ".join(parts) + "]"

    def kind_of(self, v):
        return v.typ if v.typ in ("real", "integer", "logical") else "real"

    # ---- modules ----
    def module_order(self):
        order, seen = [], set()

        def visit(m):
            if m.name in seen:
                return
            seen.add(m.name)
            for mn, _ in m.uses:
                if mn in self.program.modules:
                    visit(self.program.modules[mn])
            order.append(m)
        for m in self.program.modules.values():
            visit(m)
        return order

    def gen_module_init(self, m):
        sc = Scope(self, m, None)
        self.emit(0, f"M_{m.name} = _NS()")
        self.emit(0, f"def _init_M_{m.name}():")
        n = 0
        for v in m.vars.values():
            if v.typ == "character" or v.typ.startswith("type"):
                if v.param is not None and v.typ == "character" and v.dims is None:
                    try:
                        self.emit(1, f"M_{m.name}.{pyname(v.name)} = {parse_expr(v.param).py(sc)}")
                        n += 1
                    except Exception:
                        pass
                continue
            tgt = f"M_{m.name}.{pyname(v.name)}"
            try:
                if v.dims is not None:
                    if v.allocatable or v.pointer or any(hi is None for _, hi in v.dims):
                        self.emit(1, f"{tgt} = None")
                    else:
                        self.emit(1, f"{tgt} = FA.new('{self.kind_of(v)}', {self.dims_py(sc, v)})")
                        if v.param is not None:
                            self.emit(1, f"{tgt}.a[...] = np.reshape({parse_expr(v.param).py(sc)}, {tgt}.a.shape, order='F') "
                                         f"if np.ndim({parse_expr(v.param).py(sc)}) else {parse_expr(v.param).py(sc)}")
                else:
                    init = parse_expr(v.param).py(sc) if v.param is not None else {"integer": "0", "real": "0.0", "logical": "False"}[self.kind_of(v)]
                    if v.typ == "integer" and v.param is not None and parse_expr(v.param).typ(sc) == "real":
                        init = f"int({init})"
                    self.emit(1, f"{tgt} = {init}")
                n += 1
            except Exception as e:   # an entity this translator cannot initialise: fail where it is used, not here
                self.emit(1, f"{tgt} = None  # {type(e).__name__}: {e}")
                n += 1
        for no, s in getattr(m, "data", []):
            n += self.gen_data(sc, 1, s)
        for x, y in getattr(m, "equiv", []):   # y becomes a view of x's storage (both start at their first element)
            self.emit(1, f"M_{m.name}.{pyname(y)}.a = M_{m.name}.{pyname(x)}.a.reshape(-1, order='F')[:M_{m.name}.{pyname(y)}.a.size]"
                         f".reshape(M_{m.name}.{pyname(y)}.a.shape, order='F')")
            n += 1
        if n == 0:
            self.emit(1, "pass")

    def gen_data(self, sc, ind, stmt):
        """data a /v1, v2, .../ [, b /.../]: whole variables only, r*c repeats."""
        body = stmt[4:].strip()
        n = 0
        for m in re.finditer(r"([a-z_]\w*)\s*/([^/]*)/", body):
            name, vals = m.group(1), m.group(2)
            items = []
            for it in split_top(vals):
                if "*" in it and re.match(r"\s*\d+\s*\*", it):
                    r, c = it.split("*", 1)
                    items += [parse_expr(c.strip()).py(sc)] * int(r)
                else:
                    items.append(parse_expr(it.strip()).py(sc))
            ref = sc.ref(name)
            if sc.is_array(name):
                self.emit(ind, f"{ref}.a[...] = np.reshape(np.array([{', '.join(items)}]), {ref}.a.shape, order='F')")
            else:
                self.emit(ind, f"{ref} = {items[0]}")
            n += 1
        return n

    # ---- procedures ----
    def gen_proc(self, p, ind=0):
        mod = p.mod
        sc = Scope(self, mod, p)
        args = ", ".join(f"{pyname(a)}=None" for a in p.args)
        self.emit(ind, f"def {self.proc_pyname(p)}({args}):")
        i1 = ind + 1
        self.emit(i1, f"# {p.file}:{p.line}")
        statics = [v for v in p.vars.values() if v.static]
        # nonlocal declarations of host scalars this internal procedure assigns
        if p.host is not None:
            body_text = "\n".join(s for _, s in p.body)
            nl = []
            h = p.host
            while h is not None:
                for v in h.vars.values():
                    if v.dims is None and not v.static and v.name not in p.vars and \
                            re.search(r"(^|\n|\)\s)\s*%s\s*=(?!=)" % re.escape(v.name), body_text):
                        nl.append(pyname(v.name))
                h = h.host
            if nl:
                self.emit(i1, "nonlocal " + ", ".join(sorted(set(nl))))
        # dummies
        for a in p.args:
            v = p.vars.get(a)
            if v is None:
                continue   # a dummy procedure or an undeclared dummy
            if v.dims is not None and v.pointer:
                # a pointer dummy keeps the bounds of what it is associated with (e.g. FLX(:,:,0:LM) in IRR Update);
                # an assumed-shape dummy starts at 1 whatever the actual's bounds
                self.emit(i1, f"{pyname(a)} = _bind_pointer({pyname(a)}, '{a}')")
            elif v.dims is not None:
                self.emit(i1, f"{pyname(a)} = _bind({pyname(a)}, '{self.kind_of(v)}', {self.dims_py(sc, v)}, '{a}')")
        # parameters first (dims of locals may use them), then locals
        for v in p.vars.values():
            if v.name in p.args or v.static:
                continue
            if v.is_param and v.dims is None and v.typ != "character":
                self.emit(i1, f"{pyname(v.name)} = {parse_expr(v.param).py(sc)}")
        for v in p.vars.values():
            if v.name in p.args or v.static or (v.is_param and v.dims is None):
                continue
            if v.name in p.stmt_functions:
                continue
            if v.typ == "character" or v.typ.startswith("type"):
                self.emit(i1, f"{pyname(v.name)} = None")
                continue
            if v.dims is not None:
                if v.allocatable or v.pointer or any(hi is None for _, hi in v.dims):
                    self.emit(i1, f"{pyname(v.name)} = None")
                else:
                    self.emit(i1, f"{pyname(v.name)} = FA.new('{self.kind_of(v)}', {self.dims_py(sc, v)})")
                    if v.param is not None:
                        self.emit(i1, f"{pyname(v.name)}.a[...] = np.reshape({parse_expr(v.param).py(sc)}, {pyname(v.name)}.a.shape, order='F')")
            else:
                self.emit(i1, f"{pyname(v.name)} = " + {"integer": "0", "real": "0.0", "logical": "False"}[self.kind_of(v)])
        for name, (sargs, expr) in p.stmt_functions.items():
            for a in sargs:
                sc.push_temp_int(a) if p.vars.get(a) is None or p.vars[a].typ == "integer" else None
            self.emit(i1, f"{name}__sf = lambda {', '.join(pyname(a) for a in sargs)}: {parse_expr(expr).py(sc)}")
            for a in sargs:
                if a in sc.temp_ints:
                    sc.pop_temp_int(a)
        for no, s in getattr(p, "data", []):
            self.gen_data(sc, i1, s)
        for q in p.internal.values():
            self.gen_proc(q, i1)
        self.ret_text = {}
        self.gen_body(sc, p, i1)
        self.emit(i1, self.return_stmt(p))
        self.emit(ind, "")
        if statics:
            self.emit(ind, f"S_{p.pyid} = _NS()")
            ssc = Scope(self, mod, p)
            for v in statics:
                tgt = f"S_{p.pyid}.{pyname(v.name)}"
                if v.dims is not None:
                    self.deferred_static.append((p, v))
                    self.emit(ind, f"{tgt} = None")
                else:
                    init = parse_expr(v.param).py(ssc) if v.param is not None else "0"
                    self.deferred_static.append((p, v))
                    self.emit(ind, f"{tgt} = None")

    def return_stmt(self, p):
        outs = [pyname(a) for a in p.out_scalars]
        if p.kind == "function":
            return f"return {pyname(p.result)}"
        if not outs:
            return "return None"
        return "return (" + ", ".join(outs) + ",)"

    # ---- statements ----
    def gen_body(self, sc, p, ind):
        where_stack = []
        n_emitted_at = [len(self.out)]
        block_start = []   # output length at the start of every open block (to insert `pass` into empty ones)
        self.do_names = []  # construct name (or None) of every open do, innermost last

        def open_block():
            block_start.append(len(self.out))

        def close_block(ind_):
            start = block_start.pop()
            if len(self.out) == start:
                self.emit(ind_, "pass")
        for no, s in p.body:
            sc.line = no
            try:
                ind = self.gen_stmt(sc, p, s, ind, where_stack, open_block, close_block)
            except Exception as e:
                raise type(e)(f"{p.file}:{no}: {s[:120]!r}: {e}") from e

    def gen_stmt(self, sc, p, s, ind, where_stack, open_block, close_block):
        # construct names: `name: if` is dropped; `name: do` wraps the loop body in a try block so that `cycle name` /
        # `exit name` issued from an inner loop (LW/src/rrtmg_lw_cldprmc.F90:372-377) reach the named loop
        label = None
        m = re.match(r"([a-z_]\w*)\s*:\s*(do|if)\b(.*)$", s)
        if m and not re.match(r"[a-z_]\w*\s*::", s):
            s = m.group(2) + m.group(3)
            if m.group(2) == "do":
                label = m.group(1)
        if s == "do" or re.match(r"do\s", s):
            ind1 = self._gen_do(sc, s, ind, open_block)
            self.do_names.append(label)
            if label:
                self.emit(ind1, "try:")
                open_block()
                ind1 += 1
            return ind1
        if re.match(r"end\s*do\b", s):
            label = self.do_names.pop()
            if label:
                close_block(ind)
                ind -= 1
                self.emit(ind, "except NamedCycle as _e:")
                self.emit(ind + 1, f"if _e.args[0] != {label!r}: raise")
                self.emit(ind, "except NamedExit as _e:")
                self.emit(ind + 1, f"if _e.args[0] != {label!r}: raise")
                self.emit(ind + 1, "break")
            close_block(ind)
            return ind - 1
        if s in ("continue",) or re.match(r"(write|print|format|read|open|close|flush)\b", s):
            self.emit(ind, "pass")
            return ind
        if re.match(r"(error\s+)?stop\b", s):
            msg = s.split("stop", 1)[1].strip() or "''"
            self.emit(ind, f"raise StopError({msg!r})")
            return ind
        if s == "return":
            self.emit(ind, self.return_stmt(p))
            return ind
        m = re.match(r"(cycle|exit)\b\s*([a-z_]\w*)?$", s)
        if m:
            name = m.group(2)
            if name and name not in self.do_names:
                raise SyntaxError(f"{m.group(1)} {name}: no such open do construct")
            if name is None or name == self.do_names[-1]:
                self.emit(ind, "continue" if m.group(1) == "cycle" else "break")
            else:
                self.emit(ind, f"raise {'NamedCycle' if m.group(1) == 'cycle' else 'NamedExit'}({name!r})")
            return ind
        # ---- MAPL macros (oracle/ref_recipe/MAPL_Generic.h states what they mean) ----
        m = re.match(r"_assert\s*\((.*)\)$", s)
        if m:
            cond, msg = split_top(m.group(1))[0], ",".join(split_top(m.group(1))[1:])
            self.emit(ind, f"if not ({parse_expr(cond).py(sc)}):")
            self.emit(ind + 1, f"raise StopError('_ASSERT: ' + {parse_expr(msg).py(sc) if msg else repr('')})")
            return ind
        m = re.match(r"_fail\s*\((.*)\)$", s)
        if m:
            self.emit(ind, f"raise StopError('_FAIL: ' + {parse_expr(m.group(1)).py(sc)})")
            return ind
        m = re.match(r"_return\s*\((.*)\)$", s)
        if m:
            if "rc" in p.vars:
                self.emit(ind, "rc = 0")
            self.emit(ind, self.return_stmt(p))
            return ind
        m = re.match(r"_verify\s*\((.*)\)$", s)
        if m:
            self.emit(ind, f"if ({parse_expr(m.group(1)).py(sc)}) != 0:")
            if "rc" in p.vars:
                self.emit(ind + 1, f"rc = {parse_expr(m.group(1)).py(sc)}")
            self.emit(ind + 1, self.return_stmt(p))
            return ind
        # ---- if ----
        m = re.match(r"if\s*\((.*)\)\s*then$", s)
        if m and _balanced(m.group(1)):
            self.emit(ind, f"if {parse_expr(m.group(1)).py(sc)}:")
            open_block()
            return ind + 1
        m = re.match(r"else\s*if\s*\((.*)\)\s*then$", s)
        if m:
            close_block(ind)
            self.emit(ind - 1, f"elif {parse_expr(m.group(1)).py(sc)}:")
            open_block()
            return ind
        if s == "else":
            close_block(ind)
            self.emit(ind - 1, "else:")
            open_block()
            return ind
        if re.match(r"end\s*if\b", s):
            close_block(ind)
            return ind - 1
        if s.startswith("if"):
            m = re.match(r"if\s*\(", s)
            if m:
                j = _match_paren(s, m.end() - 1)
                cond, rest = s[m.end():j], s[j + 1:].strip()
                self.emit(ind, f"if {parse_expr(cond).py(sc)}:")
                open_block()
                ind2 = self.gen_stmt(sc, p, rest, ind + 1, where_stack, open_block, close_block)
                close_block(ind2)
                return ind
        # ---- where ----
        m = re.match(r"where\s*\(", s)
        if m:
            j = _match_paren(s, m.end() - 1)
            mask, rest = s[m.end():j], s[j + 1:].strip()
            mv = f"_wm{len(where_stack)}"
            self.emit(ind, f"{mv} = np.asarray({parse_expr(mask).py(sc)})")
            self.emit(ind, f"{mv}_rest = np.logical_not({mv})")
            if rest:   # one-line where
                self.gen_where_assign(sc, rest, ind, mv)
                return ind
            where_stack.append(mv)
            return ind
        m = re.match(r"else\s*where\s*(\((.*)\))?$", s)
        if m and where_stack:
            mv = where_stack[-1]
            if m.group(2):
                self.emit(ind, f"{mv} = np.logical_and({mv}_rest, {parse_expr(m.group(2)).py(sc)})")
                self.emit(ind, f"{mv}_rest = np.logical_and({mv}_rest, np.logical_not({mv}))")
            else:
                self.emit(ind, f"{mv} = {mv}_rest")
            return ind
        if re.match(r"end\s*where\b", s):
            where_stack.pop()
            return ind
        if where_stack:
            self.gen_where_assign(sc, s, ind, where_stack[-1])
            return ind
        # ---- call ----
        m = re.match(r"call\s+([a-z_]\w*)\s*(\((.*)\))?$", s)
        if m:
            name = m.group(1)
            args = Parser(tokenize("(" + (m.group(3) or "") + ")"))
            args.next()
            arglist = args.arglist()
            r = sc.lookup(name)
            if r is None or r[0] != "proc":
                if name.startswith("mapl_timer"):
                    self.emit(ind, "pass")
                    return ind
                raise NameError(f"unknown subroutine {name}")
            self.emit(ind, self.call_text(sc, r[2], arglist, as_function=False))
            return ind
        # ---- allocate / deallocate / nullify ----
        m = re.match(r"allocate\s*\((.*)\)$", s)
        if m:
            for it in split_top(m.group(1)):
                if re.match(r"(stat|errmsg|source|mold)\s*=", it):
                    continue
                mm = re.match(r"([a-z_]\w*)\s*\((.*)\)$", it)
                v = sc.var(mm.group(1))
                dims = parse_dims(mm.group(2))
                parts = []
                for lo, hi in dims:
                    parts.append(f"({parse_expr(lo).py(sc) if lo else '1'}, {parse_expr(hi).py(sc)})")
                self.emit(ind, f"{sc.ref(mm.group(1))} = FA.new('{self.kind_of(v)}', [{', '.join(parts)}])")
            return ind
        m = re.match(r"(deallocate|nullify)\s*\((.*)\)$", s)
        if m:
            for it in split_top(m.group(2)):
                if re.match(r"(stat|errmsg)\s*=", it):
                    continue
                self.emit(ind, f"{sc.ref(it.strip())} = None")
            return ind
        # ---- pointer assignment ----
        if "=>" in s and not s.startswith("use"):
            l, r = [x.strip() for x in s.split("=>", 1)]
            re_ = parse_expr(r)
            if isinstance(re_, Name):
                self.emit(ind, f"{sc.ref(l)} = {sc.ref(re_.name)}")
            elif isinstance(re_, Apply) and re_.name == "null":
                self.emit(ind, f"{sc.ref(l)} = None")
            else:
                self.emit(ind, f"{sc.ref(l)} = FA({re_.py(sc)})")
            return ind
        # ---- assignment ----
        k = _find_assign(s)
        if k > 0:
            self.gen_assign(sc, p, s[:k].strip(), s[k + 1:].strip(), ind)
            return ind
        raise SyntaxError("statement not understood")

    def _gen_do(self, sc, s, ind, open_block):
        m = re.match(r"do\s+([a-z_]\w*)\s*=\s*(.*)$", s)
        if m:
            parts = split_top(m.group(2))
            var = sc.ref(m.group(1))
            lo, hi = parse_expr(parts[0]).py(sc), parse_expr(parts[1]).py(sc)
            if len(parts) == 3:
                st = parse_expr(parts[2]).py(sc)
                if re.fullmatch(r"\(?-\d+\)?", st):
                    self.emit(ind, f"for {var} in range({lo}, ({hi})-1, {st}):")
                elif re.fullmatch(r"\d+", st):
                    self.emit(ind, f"for {var} in range({lo}, ({hi})+1, {st}):")
                else:
                    self.emit(ind, f"for {var} in (range({lo}, ({hi})+1, {st}) if ({st}) > 0 else range({lo}, ({hi})-1, {st})):")
            else:
                self.emit(ind, f"for {var} in range({lo}, ({hi})+1):")
            open_block()
            return ind + 1
        if s == "do":
            self.emit(ind, "while True:")
            open_block()
            return ind + 1
        m = re.match(r"do\s+while\s*\((.*)\)$", s)
        if m:
            self.emit(ind, f"while {parse_expr(m.group(1)).py(sc)}:")
            open_block()
            return ind + 1
        raise SyntaxError("do statement not understood")

    def gen_where_assign(self, sc, s, ind, mv):
        lhs, rhs = s.split("=", 1)
        l = parse_expr(lhs.strip())
        r = parse_expr(rhs.strip())
        tgt = l.py(sc) if isinstance(l, Apply) else sc.ref(l.name) + ".a"
        self.emit(ind, f"np.copyto({tgt}, {r.py(sc)}, where={mv})")

    def gen_assign(self, sc, p, lhs, rhs, ind):
        l = parse_expr(lhs)
        r = parse_expr(rhs)
        rt = r.typ(sc)
        rp = r.py(sc)
        if isinstance(l, Name):
            lt = sc.type_of(l.name)
            if sc.is_array(l.name):
                ref = sc.ref(l.name)
                if r.rank(sc) > 0:
                    self.emit(ind, f"{ref}.a[...] = np.reshape({rp}, {ref}.a.shape, order='F') if np.size({rp}) == {ref}.a.size and np.ndim({rp}) != {ref}.a.ndim else {rp}")
                else:
                    self.emit(ind, f"{ref}.a[...] = {rp}")
                return
            if lt == "integer":
                if rt == "real":
                    rp = f"int({rp})"
                elif _needs_wrap(r):
                    rp = f"_i32({rp})"
            elif lt == "real" and rt == "integer":
                rp = f"float({rp})"
            self.emit(ind, f"{sc.ref(l.name)} = {rp}")
            return
        if isinstance(l, Apply):
            lt = sc.type_of(l.name)
            if lt == "integer" and rt == "real" and r.rank(sc) == 0:
                rp = f"int({rp})"
            self.emit(ind, f"{l.index_py(sc)} = {rp}")
            return
        raise SyntaxError("bad assignment target")

    # ---- calls ----
    def call_text(self, sc, callee, actuals, as_function):
        pos = [a for a in actuals if not isinstance(a, Kw)]
        kws = {a.name: a.val for a in actuals if isinstance(a, Kw)}
        bound = {}
        for d, a in zip(callee.args, pos):
            bound[d] = a
        for k, a in kws.items():
            if k not in callee.args:
                raise NameError(f"{callee.name} has no dummy {k}")
            bound[k] = a
        parts, backs = [], []
        for d in callee.args:
            if d not in bound:
                continue
            a = bound[d]
            dv = callee.vars.get(d)
            if dv is not None and dv.dims is not None:
                parts.append(f"{pyname(d)}={self.array_actual(sc, a)}")
            else:
                parts.append(f"{pyname(d)}={a.py(sc) if not (isinstance(a, Name) and sc.lookup(a.name) and sc.lookup(a.name)[0] == 'proc') else sc.ref(a.name)}")
        call = f"{self.proc_pyname(callee)}({', '.join(parts)})"
        if as_function:
            return call
        targets = []
        for d in callee.out_scalars:
            a = bound.get(d)
            if a is None:
                targets.append("_")
            elif isinstance(a, Name) and not sc.is_array(a.name):
                targets.append(sc.ref(a.name))
            elif isinstance(a, Apply) and sc.is_array(a.name) and a.rank(sc) == 0:
                targets.append(a.index_py(sc))
            else:
                targets.append("_")
        if targets and any(t != "_" for t in targets):
            return f"{', '.join(targets)}, = {call}" if len(targets) == 1 else f"{', '.join(targets)} = {call}"
        return call

    def array_actual(self, sc, a):
        if isinstance(a, Name):
            return sc.ref(a.name)                       # the FA itself (or None for an absent optional)
        if isinstance(a, Apply) and sc.is_array(a.name):
            if a.rank(sc) > 0:
                return f"FA({a.index_py(sc)})"          # a section: a view
            # an element: sequence association from that element on
            ref = sc.ref(a.name)
            lbs = sc.lower_bounds(a.name)
            idx = ", ".join(f"({x.py(sc)})-{lbs[d] if lbs[d] is not None else f'{ref}.lb[{d}]'}" for d, x in enumerate(a.args))
            return f"FA({ref}.a.reshape(-1, order='F')[np.ravel_multi_index(({idx},), {ref}.a.shape, order='F'):])"
        return f"FA(np.asarray({a.py(sc)}))"            # an expression: a temporary

    # ---- whole program ----
    def generate(self):
        self.out = ["# generated by oracle/refexec/f90py.py - do not edit", "from types import SimpleNamespace as _NS"]
        self.deferred_static = []
        order = self.module_order()
        for m in order:
            self.gen_module_init(m)
        for m in order:
            for p in m.procs.values():
                self.gen_proc(p)
        for p in self.program.loose_procs.values():
            self.gen_proc(p)
        # static (initialised / saved) locals
        self.emit(0, "def _init_statics():")
        n = 0
        for p, v in self.deferred_static:
            sc = Scope(self, p.mod, p)
            tgt = f"S_{p.pyid}.{pyname(v.name)}"
            if v.dims is not None:
                self.emit(1, f"{tgt} = FA.new('{self.kind_of(v)}', {self.dims_py(sc, v)})")
                if v.param is not None:
                    self.emit(1, f"{tgt}.a[...] = np.reshape({parse_expr(v.param).py(sc)}, {tgt}.a.shape, order='F')")
            else:
                init = parse_expr(v.param).py(sc) if v.param is not None else {"integer": "0", "real": "0.0", "logical": "False"}.get(self.kind_of(v), "None")
                self.emit(1, f"{tgt} = {init}")
            n += 1
        if n == 0:
            self.emit(1, "pass")
        self.emit(0, "def _init_all():")
        for m in order:
            self.emit(1, f"_init_M_{m.name}()")
        self.emit(1, "_init_statics()")
        return "\n".join(self.out) + "\n"


def _needs_wrap(node):
    """Integer expressions that can leave the 32-bit range: products and shifts."""
    if isinstance(node, Bin):
        return node.op in ("*", "**") or _needs_wrap(node.a) or _needs_wrap(node.b)
    if isinstance(node, Un):
        return _needs_wrap(node.x)
    if isinstance(node, Apply):
        return node.name in ("ishft", "ieor") or any(_needs_wrap(a) for a in node.args if isinstance(a, Node) and not isinstance(a, (Slice, Kw)))
    return False


def _find_assign(s):
    """Index of the `=` of an assignment statement (outside parentheses and strings, not part of == /= <= >= =>), or -1."""
    depth, q = 0, None
    for j, ch in enumerate(s):
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
            prev = s[j - 1] if j else ""
            nxt = s[j + 1] if j + 1 < len(s) else ""
            if prev in "=/<>" or nxt in "=>":
                continue
            return j
    return -1


def _match_paren(s, i):
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
    raise SyntaxError("unbalanced parentheses")


def _balanced(s):
    depth = 0
    for ch in s:
        depth += ch == "("
        depth -= ch == ")"
        if depth < 0:
            return False
    return depth == 0


def load(files, defines=()):
    """Translate the given Fortran files; returns the namespace of the generated module (after _init_all())."""
    prog = Program()
    for f in files:
        prog.add_source(open(f, errors="replace").read(), f, defines)
    src = Translator(prog).generate()
    ns = dict(RUNTIME)
    exec(compile(src, "<f90py>", "exec"), ns)
    ns["_init_all"]()
    ns["__source__"] = src
    return ns
