"""A minimal Scheme interpreter that EXECUTES the reference's own .scm sources.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): the product path never imports it.

Why it exists: the reference is Gauche Scheme and no Scheme runtime is installed in the build image
(SURVEY.md §8c), so the reference could not be run to produce golden vectors.  This module
implements just enough of R7RS + the Gauche extensions the reference uses (modules with `use
:prefix`, `define-inline`, `receive` / `let-values`, `dotimes`, `inc!`, `push!`, generalised
`set!`, `let-optionals*`, hygienic non-ellipsis `syntax-rules`, `f64vector-*`, `gauche.array`
`array-mul`, srfi-27 `random-real` supplied by the caller) to load `/root/reference/*.scm`
UNMODIFIED and call their procedures.  `tests/golden/make_reference_golden.py` and
`make_reference_render.py` use it, in the build container only, to freeze reference outputs into
`tests/golden/ref_*`; the oracle and the CUDA path are then checked against those files.  Nothing is
copied from the reference: the sources are read from where they lie at generation time.  It also
executes the repo's own Gauche host modules (`scheme_raytrace_b200/scheme/*.scm`, which add `format`
directives, file ports, a few srfi-1 procedures to the vocabulary) in `tests/test_scheme_host.py`;
its own semantics are pinned by `tests/test_minischeme.py`.

Fidelity notes (where Gauche's behaviour had to be restated rather than executed):
  * numbers: exact integers / rationals (fractions.Fraction) / IEEE doubles with the usual
    contagion; `(/ 1 0.0)` = +inf.0; multi-argument `+ * - /` fold left to right; `sqrt` of an exact
    square stays exact; `sqrt` / `asin` outside their real domain return NaN (Gauche: a complex);
  * libm functions come from Python's `math` (the platform libm, as Gauche's do);
  * `f64vector-dot` accumulates in index order without FMA; `array-mul` accumulates
    sum_k a[i][k] * b[k][j] from exact 0 in index order (lib/gauche/array.scm);
  * `sort` is a stable merge sort on the caller's predicate (the reference's comparators return
    -1 / 1, both true - SURVEY Q12 - so its BVH builders are not used for golden vectors);
  * a multiple-value result used as a procedure ARGUMENT keeps its first value (Gauche's behaviour; R7RS leaves it
    undefined) - `get-normal` of the Klein primitive relies on it;
  * procedure arguments are evaluated left to right (R7RS and Gauche leave the order unspecified).  It only decides WHICH
    draw feeds which component where one call has several `(random-real)` operands (util.scm:10,18; the three evaluations
    of Q15): the draws are i.i.d., so distributions, and therefore every parity statement, are unaffected;
  * `random-real` is whatever callable the host installs: random STREAMS are not a parity target.
"""
import math
import os
from fractions import Fraction


class Sym(str):
    __slots__ = ()


class Keyword(str):
    __slots__ = ()


class Alias:
    """A template identifier of a macro expansion: resolves in the macro's definition module unless
    the same expansion bound it (hygiene by renaming)."""
    __slots__ = ("name", "menv", "eid")

    def __init__(self, name, menv, eid):
        self.name, self.menv, self.eid = name, menv, eid

    def __hash__(self):
        return hash((self.name, self.eid))

    def __eq__(self, o):
        return isinstance(o, Alias) and o.name == self.name and o.eid == self.eid

    def __repr__(self):
        return f"#<alias {self.name}>"


class Nil:
    __slots__ = ()

    def __repr__(self):
        return "()"

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0


NIL = Nil()


class Pair:
    __slots__ = ("car", "cdr")

    def __init__(self, car, cdr):
        self.car, self.cdr = car, cdr

    def __iter__(self):
        p = self
        while isinstance(p, Pair):
            yield p.car
            p = p.cdr

    def __len__(self):
        return sum(1 for _ in self)

    def __repr__(self):
        return "(" + " ".join(repr(x) for x in self) + ")"


def to_list(seq):
    out = NIL
    for x in reversed(list(seq)):
        out = Pair(x, out)
    return out


class F64(list):
    """f64vector"""
    __slots__ = ()


class Values(tuple):
    __slots__ = ()


class Array:
    def __init__(self, rows, cols, data):
        self.rows, self.cols, self.data = rows, cols, data


class SchemeError(Exception):
    pass


def scheme_format(*a):
    """(format [#f] "control string" arg ...) -> string, with the directives the reference uses: ~D ~A ~S ~% ~~
    (main.scm:442,449 builds the PPM header and rows with ~D)."""
    if a and a[0] is False:
        a = a[1:]
    if not a or not isinstance(a[0], str) or isinstance(a[0], Sym):
        return " ".join(map(str, a))
    fmt, args, out, i, k = a[0], list(a[1:]), [], 0, 0
    while i < len(fmt):
        c = fmt[i]
        if c == "~" and i + 1 < len(fmt):
            dct = fmt[i + 1].upper()
            i += 2
            if dct in "DAS":
                out.append(str(args[k])); k += 1
            elif dct == "%":
                out.append("\n")
            elif dct == "~":
                out.append("~")
            else:
                raise SchemeError(f"format: unsupported directive ~{dct}")
        else:
            out.append(c); i += 1
    return "".join(out)


class _Eof:
    def __repr__(self):
        return "#<eof>"


EOF = _Eof()


DOT = Sym(".")


# ------------------------------------------------------------------------------------------------
# reader
def tokenize(src):
    i, n, out = 0, len(src), []
    while i < n:
        c = src[i]
        if c.isspace():
            i += 1
        elif c == ";":
            while i < n and src[i] != "\n":
                i += 1
        elif src.startswith("#|", i):
            depth, i = 1, i + 2
            while depth and i < n:
                if src.startswith("#|", i):
                    depth, i = depth + 1, i + 2
                elif src.startswith("|#", i):
                    depth, i = depth - 1, i + 2
                else:
                    i += 1
        elif src.startswith("#;", i):
            out.append("#;")
            i += 2
        elif c in "()[]":
            out.append("(" if c in "([" else ")")
            i += 1
        elif c in "'`":
            out.append(c)
            i += 1
        elif c == ",":
            if src.startswith(",@", i):
                out.append(",@")
                i += 2
            else:
                out.append(",")
                i += 1
        elif c == '"':
            j, buf = i + 1, []
            while src[j] != '"':
                if src[j] == "\\":
                    j += 1
                    buf.append({"n": "\n", "t": "\t", "\\": "\\", '"': '"'}.get(src[j], src[j]))
                else:
                    buf.append(src[j])
                j += 1
            out.append(("str", "".join(buf)))
            i = j + 1
        else:
            j = i
            while j < n and not src[j].isspace() and src[j] not in "()[]\";'":
                j += 1
            out.append(("atom", src[i:j]))
            i = j
    return out


def parse_atom(tok):
    if tok in ("#t", "#true"):
        return True
    if tok in ("#f", "#false"):
        return False
    if tok in ("+inf.0", "-inf.0"):
        return float(tok[:4])
    if tok == "+nan.0":
        return float("nan")
    try:
        return int(tok)
    except ValueError:
        pass
    if "/" in tok:
        a, _, b = tok.partition("/")
        if a.lstrip("+-").isdigit() and b.isdigit():
            return norm(Fraction(int(a), int(b)))
    try:
        if tok[0].isdigit() or (tok[0] in "+-." and len(tok) > 1 and (tok[1].isdigit() or tok[1] == ".")):
            return float(tok)
    except ValueError:
        pass
    if tok.startswith("#\\"):
        return tok[2:]
    if tok.startswith(":") and len(tok) > 1:
        return Keyword(tok)
    return Sym(tok)


def read_all(src):
    toks = tokenize(src)
    pos = 0

    def read():
        nonlocal pos
        t = toks[pos]
        pos += 1
        if t == "(":
            lst = []
            while toks[pos] != ")":
                lst.append(read())
            pos += 1
            return [x for x in lst if x is not _SKIP]
        if t == ")":
            raise SchemeError("unexpected )")
        if t == "'":
            return [Sym("quote"), read()]
        if t == "`":
            return [Sym("quasiquote"), read()]
        if t == ",":
            return [Sym("unquote"), read()]
        if t == ",@":
            return [Sym("unquote-splicing"), read()]
        if t == "#;":
            read()
            return _SKIP
        kind, val = t
        return val if kind == "str" else parse_atom(val)

    forms = []
    while pos < len(toks):
        f = read()
        if f is not _SKIP:
            forms.append(f)
    return forms


_SKIP = object()


# ------------------------------------------------------------------------------------------------
# numbers
def norm(x):
    if isinstance(x, Fraction) and x.denominator == 1:
        return int(x)
    return x


def is_exact(x):
    return isinstance(x, (int, Fraction)) and not isinstance(x, bool)


def n_add(*a):
    r = 0
    for x in a:
        r = r + x
    return norm(r)


def n_mul(*a):
    r = 1
    for x in a:
        r = r * x
    return norm(r)


def n_sub(a, *rest):
    if not rest:
        return -a
    for x in rest:
        a = a - x
    return norm(a)


def _div2(a, b):
    if is_exact(a) and is_exact(b):
        if b == 0:
            raise SchemeError("attempt to calculate a division by zero")
        return norm(Fraction(a, b))
    a, b = float(a), float(b)
    if b == 0.0:
        if a == 0.0 or a != a:
            return float("nan")
        return math.copysign(float("inf"), a) * math.copysign(1.0, b)
    return a / b


def n_div(a, *rest):
    if not rest:
        return _div2(1, a)
    for x in rest:
        a = _div2(a, x)
    return a


def n_cmp(op):
    def f(*a):
        return all(op(a[i], a[i + 1]) for i in range(len(a) - 1))
    return f


def n_sqrt(x):
    if is_exact(x) and x >= 0:
        if isinstance(x, int):
            r = math.isqrt(x)
            if r * r == x:
                return r
        else:
            rn, rd = math.isqrt(x.numerator), math.isqrt(x.denominator)
            if rn * rn == x.numerator and rd * rd == x.denominator:
                return norm(Fraction(rn, rd))
    x = float(x)
    return math.sqrt(x) if x >= 0 else float("nan")


def n_minmax(fn):
    def f(*a):
        r = a[0]
        for x in a[1:]:
            r = x if (fn is min and x < r) or (fn is max and x > r) else r
        return float(r) if any(isinstance(x, float) for x in a) else r
    return f


def n_log(x, base=None):
    def ln(v):
        v = float(v)
        if v == 0.0:
            return float("-inf")
        return math.log(v) if v > 0 else float("nan")
    if is_exact(x) and x == 1 and base is None:
        return 0
    return ln(x) if base is None else ln(x) / ln(base)


def n_expt(a, b):
    if is_exact(a) and isinstance(b, int):
        return norm(Fraction(a) ** b)
    return math.pow(float(a), float(b))


def n_asin(x):
    x = float(x)
    return math.asin(x) if -1.0 <= x <= 1.0 else float("nan")


def n_acos(x):
    x = float(x)
    return math.acos(x) if -1.0 <= x <= 1.0 else float("nan")


def n_atan(y, x=None):
    return math.atan(float(y)) if x is None else math.atan2(float(y), float(x))


def floor_exact(x):
    return int(math.floor(x))


def n_exact(x):
    if isinstance(x, float):
        return norm(Fraction(x))
    return x


def n_clamp(x, lo=None, hi=None):
    r = x
    if lo is not None and r < lo:
        r = lo
    if hi is not None and r > hi:
        r = hi
    return float(r) if isinstance(x, float) or isinstance(r, float) else r


# ------------------------------------------------------------------------------------------------
# f64vector / array
def _f64_op(op):
    def f(a, b):
        if isinstance(b, (list, F64)):
            return F64(op(x, float(y)) for x, y in zip(a, b))
        b = float(b)
        return F64(op(x, b) for x in a)
    return f


def f64_div(a, b):
    if isinstance(b, (list, F64)):
        return F64(_div2(x, float(y)) for x, y in zip(a, b))
    return F64(_div2(x, float(b)) for x in a)


def f64_dot(a, b):
    r = 0.0
    for x, y in zip(a, b):
        r += x * y
    return r


def make_array(shape, *vals):
    r0, r1, c0, c1 = shape
    return Array(r1 - r0, c1 - c0, list(vals))


def array_mul(a, b):
    out = []
    for i in range(a.rows):
        for j in range(b.cols):
            tmp = 0
            for k in range(a.cols):
                tmp = n_add(tmp, n_mul(a.data[i * a.cols + k], b.data[k * b.cols + j]))
            out.append(tmp)
    return Array(a.rows, b.cols, out)


# ------------------------------------------------------------------------------------------------
# environments
class Env:
    __slots__ = ("vars", "parent")

    def __init__(self, parent, vars=None):
        self.vars, self.parent = ({} if vars is None else vars), parent


class Module:
    def __init__(self, name, interp):
        self.name, self.vars, self.imports, self.exports, self.export_all, self.interp = name, {}, [], set(), False, interp

    def lookup(self, sym):
        if sym in self.vars:
            return self.vars[sym]
        for table in reversed(self.imports):           # most recent `use` first
            if sym in table:
                return table[sym]
        g = self.interp.globals
        if sym in g:
            return g[sym]
        raise SchemeError(f"unbound variable: {sym} (module {self.name})")

    def exported(self):
        if self.export_all:
            return dict(self.vars)
        return {k: v for k, v in self.vars.items() if k in self.exports}


def lookup(x, env):
    e = env
    while isinstance(e, Env):
        if x in e.vars:
            return e.vars[x]
        e = e.parent
    if isinstance(x, Alias):
        return x.menv.lookup(x.name)
    return e.lookup(x)


def set_var(x, env, val):
    e = env
    while isinstance(e, Env):
        if x in e.vars:
            e.vars[x] = val
            return
        e = e.parent
    if isinstance(x, Alias):
        x.menv.vars[x.name] = val
    else:
        e.vars[x] = val


def module_of(env):
    while isinstance(env, Env):
        env = env.parent
    return env


def symname(x):
    if isinstance(x, Sym):
        return x
    if isinstance(x, Alias):
        return x.name
    return None


class Closure:
    __slots__ = ("params", "rest", "body", "env", "name")

    def __init__(self, params, rest, body, env, name=None):
        self.params, self.rest, self.body, self.env, self.name = params, rest, body, env, name


class Macro:
    def __init__(self, rules, menv):
        self.rules, self.menv = rules, menv


def parse_params(p):
    if symname(p) is not None:
        return [], p
    if DOT in p:
        i = p.index(DOT)
        return list(p[:i]), p[i + 1]
    return list(p), None


def truthy(x):
    return x is not False


# ------------------------------------------------------------------------------------------------
class Interp:
    def __init__(self, load_path, random_real=None):
        self.load_path = load_path if isinstance(load_path, (list, tuple)) else [load_path]
        self.modules = {}
        self.out = []
        self.eid = 0
        self.random_real = random_real or (lambda: 0.5)
        self.globals = {}
        self._install_builtins()
        self.user = Module("user", self)
        self.stub_modules = {"srfi-11", "srfi-13", "srfi-27", "srfi-43", "math.const", "gauche.uvector", "gauche.record", "gauche.sequence",
                             "gauche.collection", "gauche.array", "gauche.threads", "gauche.time", "gl", "gl.glut"}

    # -- module loading -----------------------------------------------------------------------
    def require(self, name):
        if name in self.modules:
            return self.modules[name]
        if name in self.stub_modules:
            m = self.modules[name] = Module(name, self)
            return m
        for d in self.load_path:
            path = os.path.join(d, name + ".scm")
            if os.path.exists(path):
                self.load_file(path)
                return self.modules[name]
        raise SchemeError(f"cannot find module {name}")

    def load_file(self, path, only=None, cur=None):
        """Evaluate the top-level forms of a file.  `only`: a predicate on the form (used to take
        selected definitions out of main.scm, which also holds the GLUT viewer)."""
        with open(path) as f:
            forms = read_all(f.read())
        cur = cur or self.user
        for form in forms:
            if isinstance(form, list) and form and form[0] == "select-module":
                cur = self.modules[form[1]]
                continue
            if only is not None and not only(form):
                continue
            self.eval(form, cur)
        return cur

    def use(self, cur, name, prefix=None):
        m = self.require(name)
        table = m.exported()
        if prefix:
            table = {Sym(prefix + k): v for k, v in table.items()}
        cur.imports.append(table)

    # -- evaluation ---------------------------------------------------------------------------
    def apply(self, f, args):
        if isinstance(f, Closure):
            return self.eval_body(f.body, self.bind(f, args))
        return f(*args)

    def bind(self, f, args):
        n = len(f.params)
        if len(args) < n or (f.rest is None and len(args) != n):
            raise SchemeError(f"wrong number of arguments for {f.name or 'lambda'}: required {n}, got {len(args)}")
        vars = dict(zip(f.params, args))
        if f.rest is not None:
            vars[f.rest] = to_list(args[n:])
        return Env(f.env, vars)

    def eval_body(self, body, env):
        for x in body[:-1]:
            self.eval(x, env)
        return self.eval(body[-1], env)

    def eval(self, x, env):
        while True:
            if isinstance(x, (Sym, Alias)):
                return lookup(x, env)
            if not isinstance(x, list):
                return x
            if not x:
                return NIL
            head = x[0]
            name = symname(head)
            if name is not None:
                sf = SPECIAL.get(name)
                if sf is not None:
                    r = sf(self, x, env)
                    if type(r) is _Tail:
                        x, env = r.x, r.env
                        continue
                    return r
                f = lookup(head, env)
                if isinstance(f, Macro):
                    x = self.expand(f, x)
                    continue
            else:
                f = self.eval(head, env)
            args = [self.eval(a, env) for a in x[1:]]
            for i, a in enumerate(args):                     # Gauche: multiple values reaching a one-value continuation keep the first
                if type(a) is Values:                        # (geometry.scm:628-633 subtracts two (dist-func ...) results, each (values d n))
                    args[i] = a[0] if a else None
            if isinstance(f, Closure):
                env = self.bind(f, args)
                for b in f.body[:-1]:
                    self.eval(b, env)
                x = f.body[-1]
                continue
            if not callable(f):
                raise SchemeError(f"invalid application: {f!r}")
            return f(*args)

    # -- syntax-rules (no ellipsis) ------------------------------------------------------------
    def expand(self, mac, form):
        for pattern, template in mac.rules:
            b = {}
            if self._match(pattern[1:], form[1:], b):
                self.eid += 1
                return self._instantiate(template, b, mac.menv, self.eid)
        raise SchemeError(f"malformed macro use: {form}")

    def _match(self, pat, form, b):
        if symname(pat) is not None:
            b[pat] = form
            return True
        if isinstance(pat, list):
            if not isinstance(form, list) or len(pat) != len(form):
                return False
            return all(self._match(p, f, b) for p, f in zip(pat, form))
        return pat == form

    def _instantiate(self, t, b, menv, eid):
        if isinstance(t, Sym):
            if t in b:
                return b[t]
            return Alias(t, menv, eid)
        if isinstance(t, list):
            return [self._instantiate(y, b, menv, eid) for y in t]
        return t

    # -- builtins -----------------------------------------------------------------------------
    def _install_builtins(self):
        g = self.globals
        import operator as op

        def d(name, fn):
            g[Sym(name)] = fn

        d("+", n_add); d("*", n_mul); d("-", n_sub); d("/", n_div)
        d("=", n_cmp(op.eq)); d("<", n_cmp(op.lt)); d(">", n_cmp(op.gt)); d("<=", n_cmp(op.le)); d(">=", n_cmp(op.ge))
        d("sqrt", n_sqrt); d("abs", abs); d("min", n_minmax(min)); d("max", n_minmax(max))
        d("sin", lambda x: math.sin(float(x))); d("cos", lambda x: math.cos(float(x))); d("tan", lambda x: math.tan(float(x)))
        d("asin", n_asin); d("acos", n_acos); d("atan", n_atan); d("exp", lambda x: math.exp(float(x))); d("log", n_log); d("expt", n_expt)
        d("floor", lambda x: float(math.floor(x)) if isinstance(x, float) else math.floor(x))
        d("ceiling", lambda x: float(math.ceil(x)) if isinstance(x, float) else math.ceil(x))
        d("round", lambda x: float(round(x)) if isinstance(x, float) else round(x))
        d("truncate", lambda x: float(math.trunc(x)) if isinstance(x, float) else math.trunc(x))
        d("floor->exact", floor_exact); d("ceiling->exact", lambda x: int(math.ceil(x))); d("round->exact", lambda x: int(round(x)))
        d("exact->inexact", float); d("inexact", float); d("exact", n_exact); d("inexact->exact", n_exact)
        d("number?", lambda x: isinstance(x, (int, float, Fraction)) and not isinstance(x, bool))
        d("zero?", lambda x: x == 0); d("positive?", lambda x: x > 0); d("negative?", lambda x: x < 0)
        d("even?", lambda x: x % 2 == 0); d("odd?", lambda x: x % 2 == 1)
        d("quotient", lambda a, b: int(a / b) if isinstance(a, float) or isinstance(b, float) else (abs(a) // abs(b)) * (1 if (a >= 0) == (b >= 0) else -1))
        d("remainder", lambda a, b: math.fmod(a, b) if isinstance(a, float) or isinstance(b, float) else a - b * ((abs(a) // abs(b)) * (1 if (a >= 0) == (b >= 0) else -1)))
        d("modulo", lambda a, b: a % b); d("fmod", lambda a, b: math.fmod(float(a), float(b)))
        d("logand", lambda *a: __import__("functools").reduce(op.and_, a, -1)); d("logior", lambda *a: __import__("functools").reduce(op.or_, a, 0))
        d("logxor", lambda *a: __import__("functools").reduce(op.xor, a, 0)); d("ash", lambda a, b: a << b if b >= 0 else a >> -b)
        d("clamp", n_clamp); d("square", lambda x: n_mul(x, x))
        d("pi", math.pi); d("pi/2", math.pi / 2); d("pi/4", math.pi / 4); d("pi/180", math.pi / 180); d("180/pi", 180 / math.pi); d("1/pi", 1 / math.pi); d("e", math.e)
        d("random-real", lambda: self.random_real())
        d("not", lambda x: x is False); d("eq?", lambda a, b: a is b or (type(a) is type(b) and isinstance(a, (int, str)) and a == b))
        d("eqv?", lambda a, b: a is b or (type(a) is type(b) and isinstance(a, (int, float, str, Fraction)) and a == b))
        d("equal?", lambda a, b: a == b if not isinstance(a, Pair) else list(a) == list(b))
        d("boolean?", lambda x: isinstance(x, bool)); d("symbol?", lambda x: isinstance(x, Sym)); d("string?", lambda x: isinstance(x, str) and not isinstance(x, (Sym, Keyword)))
        d("procedure?", lambda x: callable(x) or isinstance(x, Closure))
        # pairs and lists
        d("cons", Pair); d("car", lambda p: p.car); d("cdr", lambda p: p.cdr); d("cadr", lambda p: p.cdr.car); d("cddr", lambda p: p.cdr.cdr)
        d("caddr", lambda p: p.cdr.cdr.car); d("list", lambda *a: to_list(a)); d("null?", lambda x: x is NIL); d("pair?", lambda x: isinstance(x, Pair))
        d("list?", lambda x: x is NIL or isinstance(x, Pair)); d("length", lambda x: len(x)); d("reverse", lambda x: to_list(reversed(list(x))))
        d("reverse!", lambda x: to_list(reversed(list(x)))); d("append", lambda *a: to_list([y for x in a for y in x])); d("append!", lambda *a: to_list([y for x in a for y in x]))
        d("list-copy", lambda x: to_list(list(x))); d("last", lambda x: list(x)[-1]); d("list-ref", lambda x, i: list(x)[i]); d("list-tail", lambda x, k: to_list(list(x)[k:]))
        d("drop-right!", lambda x, k: to_list(list(x)[:len(x) - k])); d("drop-right", lambda x, k: to_list(list(x)[:len(x) - k])); d("take", lambda x, k: to_list(list(x)[:k])); d("drop", lambda x, k: to_list(list(x)[k:]))
        d("list->vector", lambda x: list(x)); d("vector->list", lambda v: to_list(v)); d("iota", lambda n, s=0, st=1: to_list([s + i * st for i in range(n)]))
        d("map", lambda f, *ls: to_list([self.apply(f, list(a)) for a in zip(*[list(l) for l in ls])]))
        d("for-each", lambda f, *ls: [self.apply(f, list(a)) for a in zip(*[list(l) for l in ls])] and None)
        d("filter", lambda f, l: to_list([x for x in l if truthy(self.apply(f, [x]))]))
        d("fold", lambda f, init, l: __import__("functools").reduce(lambda acc, x: self.apply(f, [x, acc]), list(l), init))
        d("reduce", self._reduce); d("reduce-right", self._reduce_right); d("apply", lambda f, *a: self.apply(f, list(a[:-1]) + list(a[-1])))
        d("sort", self._sort); d("subseq", lambda s, a=0, b=None: (to_list(list(s)[a:b]) if not isinstance(s, list) else s[a:b]))
        d("ref", lambda s, i: s[i] if isinstance(s, list) else list(s)[i])
        # vectors
        d("vector", lambda *a: list(a)); d("make-vector", lambda n, fill=None: [fill] * n); d("vector-ref", lambda v, i: v[i]); d("vector-set!", self._vset)
        d("vector-length", len); d("vector-tabulate", lambda n_or_f, f_or_n: [self.apply(f_or_n, [i]) for i in range(n_or_f)] if isinstance(n_or_f, int) else [self.apply(n_or_f, [i]) for i in range(f_or_n)])
        d("vector-swap!", self._vswap); d("vector-fill!", lambda v, x: v.__setitem__(slice(None), [x] * len(v))); d("vector-copy", lambda v: list(v))
        d("vector-map", lambda f, v: [self.apply(f, [x]) for x in v]); d("vector-for-each", lambda f, v: [self.apply(f, [x]) for x in v] and None)
        # uniform vectors
        d("f64vector", lambda *a: F64(float(x) for x in a)); d("make-f64vector", lambda n, fill=0.0: F64([float(fill)] * n))
        d("f64vector-ref", lambda v, i: v[i]); d("f64vector-set!", lambda v, i, x: v.__setitem__(i, float(x))); d("f64vector-length", len)
        d("f64vector-add", _f64_op(op.add)); d("f64vector-sub", _f64_op(op.sub)); d("f64vector-mul", _f64_op(op.mul)); d("f64vector-div", f64_div); d("f64vector-dot", f64_dot)
        d("f64vector-copy", lambda v: F64(v)); d("f64vector->list", lambda v: to_list(v))
        d("make-u8vector", lambda n, fill=0: [fill] * n); d("u8vector-ref", lambda v, i: v[i]); d("u8vector-set!", self._vset)
        # gauche.array
        d("shape", lambda *a: list(a)); d("array", make_array); d("array-ref", lambda a, i, j: a.data[i * a.cols + j]); d("array-mul", array_mul)
        # values, output, errors
        d("values", lambda *a: a[0] if len(a) == 1 else Values(a))
        d("call-with-values", lambda prod, cons: self.apply(cons, list(v) if isinstance(v := self.apply(prod, []), Values) else [v]))
        d("display", lambda x, *_: self.out.append(str(x))); d("print", lambda *a: self.out.append(" ".join(map(str, a)) + "\n")); d("newline", lambda *_: self.out.append("\n"))
        d("format", scheme_format); d("error", self._error); d("errorf", self._error); d("undefined", lambda: None)
        d("string-append", lambda *a: "".join(a)); d("number->string", str); d("string->number", lambda s: parse_atom(s)); d("symbol->string", str)
        d("x->string", str); d("string-split", lambda s, sep: to_list(s.split(sep)))
        # what the repo's own Gauche host (scheme_raytrace_b200/scheme/*.scm) needs on top of the reference's vocabulary
        d("set-car!", lambda p, x: setattr(p, "car", x)); d("set-cdr!", lambda p, x: setattr(p, "cdr", x))
        d("caar", lambda p: p.car.car); d("cdar", lambda p: p.car.cdr); d("cdddr", lambda p: p.cdr.cdr.cdr); d("cadddr", lambda p: p.cdr.cdr.cdr.car)
        d("fold-right", lambda f, init, l: __import__("functools").reduce(lambda acc, x: self.apply(f, [x, acc]), reversed(list(l)), init))
        d("append-map", lambda f, *ls: to_list([y for a in zip(*[list(l) for l in ls]) for y in self.apply(f, list(a))]))
        d("make-list", lambda n, fill=None: to_list([fill] * n)); d("string-join", lambda l, sep=" ": sep.join(list(l)))
        d("exact?", lambda x: isinstance(x, (int, Fraction)) and not isinstance(x, bool)); d("inexact?", lambda x: isinstance(x, float))
        d("integer?", lambda x: (isinstance(x, int) and not isinstance(x, bool)) or (isinstance(x, float) and x == math.floor(x)))
        d("with-output-to-file", self._with_output_to_file)
        d("with-input-from-file", self._with_input_from_file); d("read-line", self._read_line); d("eof-object?", lambda x: x is EOF)
        d("port-for-each", self._port_for_each)          # points.scm:12-18 reads its CSV with these
        d("run-process", lambda *a, **k: self._error("run-process: no subprocesses in this interpreter"))

    def _with_input_from_file(self, path, thunk):
        with open(path) as f:
            saved, self.inp = getattr(self, "inp", None), iter(f.read().splitlines())
        try:
            return self.apply(thunk, [])
        finally:
            self.inp = saved

    def _read_line(self, *_):
        return next(self.inp, EOF)

    def _port_for_each(self, fn, reader):
        while True:
            x = self.apply(reader, [])
            if x is EOF:
                return None
            self.apply(fn, [x])

    def _with_output_to_file(self, path, thunk):
        saved, self.out = self.out, []
        try:
            self.apply(thunk, [])
            with open(path, "w") as f:
                f.write("".join(self.out))
        finally:
            self.out = saved

    def _error(self, *a):
        raise SchemeError(" ".join(map(str, a)))

    def _vset(self, v, i, x):
        v[i] = x

    def _vswap(self, v, i, j):
        v[i], v[j] = v[j], v[i]

    def _reduce(self, f, ridentity, lst):          # srfi-1: (f elem acc), left to right
        lst = list(lst)
        if not lst:
            return ridentity
        acc = lst[0]
        for x in lst[1:]:
            acc = self.apply(f, [x, acc])
        return acc

    def _reduce_right(self, f, ridentity, lst):    # (f e1 (f e2 ... en))
        lst = list(lst)
        if not lst:
            return ridentity
        acc = lst[-1]
        for x in reversed(lst[:-1]):
            acc = self.apply(f, [x, acc])
        return acc

    def _sort(self, seq, less):
        import functools
        items = list(seq)
        out = sorted(items, key=functools.cmp_to_key(lambda a, b: -1 if truthy(self.apply(less, [a, b])) else (1 if truthy(self.apply(less, [b, a])) else 0)))
        return out if isinstance(seq, list) else to_list(out)

    # -- convenience for hosts ------------------------------------------------------------------
    def call(self, module, name, *args):
        return self.apply(self.modules[module].lookup(Sym(name)), list(args))

    def get(self, module, name):
        return self.modules[module].lookup(Sym(name))


class _Tail:
    __slots__ = ("x", "env")

    def __init__(self, x, env):
        self.x, self.env = x, env


def _body_tail(it, body, env):
    for b in body[:-1]:
        it.eval(b, env)
    return _Tail(body[-1], env)


# ------------------------------------------------------------------------------------------------
# special forms: f(interp, form, env) -> value | _Tail
def sf_quote(it, x, env):
    def conv(d):
        if isinstance(d, list):
            return to_list([conv(y) for y in d])
        if isinstance(d, Alias):
            return d.name
        return d
    return conv(x[1])


def sf_if(it, x, env):
    if truthy(it.eval(x[1], env)):
        return _Tail(x[2], env)
    return _Tail(x[3], env) if len(x) > 3 else None


def sf_define(it, x, env):
    target = x[1]
    if isinstance(target, list):                         # (define (name . params) body...)
        name = target[0]
        params, rest = parse_params(target[1:])
        val = Closure(params, rest, x[2:], env, name=str(symname(name)))
    else:
        name = target
        val = it.eval(x[2], env) if len(x) > 2 else None
        if isinstance(val, Closure) and val.name is None:
            val.name = str(symname(name))
    if isinstance(env, Env):
        env.vars[name] = val
    else:
        env.vars[symname(name)] = val
    return None


def sf_lambda(it, x, env):
    params, rest = parse_params(x[1])
    return Closure(params, rest, x[2:], env)


def sf_let(it, x, env):
    if symname(x[1]) is not None:                        # named let
        name, binds, body = x[1], x[2], x[3:]
        loop_env = Env(env)
        f = Closure([b[0] for b in binds], None, body, loop_env, name=str(symname(name)))
        loop_env.vars[name] = f
        args = [it.eval(b[1], env) for b in binds]
        return _body_tail(it, body, it.bind(f, args))
    new = Env(env, {b[0]: (it.eval(b[1], env) if len(b) > 1 else None) for b in x[1]})
    return _body_tail(it, x[2:], new)


def sf_let_star(it, x, env):
    new = env
    for b in x[1]:
        new = Env(new, {b[0]: (it.eval(b[1], new) if len(b) > 1 else None)})
    return _body_tail(it, x[2:], Env(new))


def sf_letrec(it, x, env):
    new = Env(env)
    for b in x[1]:
        val = it.eval(b[1], new)
        if isinstance(val, Closure) and val.name is None:
            val.name = str(symname(b[0]))
        new.vars[b[0]] = val
    return _body_tail(it, x[2:], new)


def sf_set(it, x, env):
    target = x[1]
    if isinstance(target, list):                         # generalised set! (srfi-17): (set! (vector-ref v i) x)
        acc = symname(target[0])
        val = it.eval(x[2], env)
        args = [it.eval(a, env) for a in target[1:]]
        if acc is not None:
            f = lookup(target[0], env)
            if isinstance(f, Macro) or isinstance(f, Closure):   # define-inline accessor wrapping vector-ref: expand one level
                inner = f.body[-1] if isinstance(f, Closure) else None
                if inner is not None and isinstance(inner, list) and symname(inner[0]) in ("vector-ref", "f64vector-ref"):
                    call_env = it.bind(f, args)
                    args = [it.eval(a, call_env) for a in inner[1:]]
                    acc = symname(inner[0])
        if acc in ("vector-ref", "f64vector-ref", "ref", "u8vector-ref"):
            args[0][args[1]] = float(val) if isinstance(args[0], F64) else val
        elif acc == "car":
            args[0].car = val
        elif acc == "cdr":
            args[0].cdr = val
        else:
            raise SchemeError(f"unsupported generalised set!: {target}")
        return None
    set_var(target, env, it.eval(x[2], env))
    return None


def sf_begin(it, x, env):
    if len(x) == 1:
        return None
    return _body_tail(it, x[1:], env)


def sf_cond(it, x, env):
    for clause in x[1:]:
        if symname(clause[0]) == "else":
            return _body_tail(it, clause[1:], env)
        test = it.eval(clause[0], env)
        if truthy(test):
            if len(clause) == 1:
                return test
            if symname(clause[1]) == "=>":
                return it.apply(it.eval(clause[2], env), [test])
            return _body_tail(it, clause[1:], env)
    return None


def sf_case(it, x, env):
    key = it.eval(x[1], env)
    for clause in x[2:]:
        if symname(clause[0]) == "else" or any(key == d for d in clause[0]):
            return _body_tail(it, clause[1:], env)
    return None


def sf_and(it, x, env):
    if len(x) == 1:
        return True
    for e in x[1:-1]:
        if not truthy(it.eval(e, env)):
            return False
    return _Tail(x[-1], env)


def sf_or(it, x, env):
    if len(x) == 1:
        return False
    for e in x[1:-1]:
        v = it.eval(e, env)
        if truthy(v):
            return v
    return _Tail(x[-1], env)


def sf_when(it, x, env):
    if truthy(it.eval(x[1], env)):
        return _body_tail(it, x[2:], env)
    return None


def sf_unless(it, x, env):
    if not truthy(it.eval(x[1], env)):
        return _body_tail(it, x[2:], env)
    return None


def sf_dotimes(it, x, env):
    var, count = x[1][0], it.eval(x[1][1], env)
    for i in range(count):
        e = Env(env, {var: i})
        for b in x[2:]:
            it.eval(b, e)
    return None


def sf_do(it, x, env):
    specs, (test, *res), body = x[1], x[2], x[3:]
    e = Env(env, {s[0]: it.eval(s[1], env) for s in specs})
    while not truthy(it.eval(test, e)):
        for b in body:
            it.eval(b, e)
        e = Env(env, {s[0]: (it.eval(s[2], e) if len(s) > 2 else e.vars[s[0]]) for s in specs})
    r = None
    for b in res:
        r = it.eval(b, e)
    return r


def _bind_formals(formals, val):
    vals = list(val) if isinstance(val, Values) else [val]
    params, rest = parse_params(formals)
    if len(vals) < len(params) or (rest is None and len(vals) != len(params)):
        raise SchemeError(f"received {len(vals)} values where {len(params)} were expected")
    vars = dict(zip(params, vals))
    if rest is not None:
        vars[rest] = to_list(vals[len(params):])
    return vars


def sf_receive(it, x, env):
    return _body_tail(it, x[3:], Env(env, _bind_formals(x[1], it.eval(x[2], env))))


def sf_let_values(it, x, env):
    vars = {}
    for formals, expr in x[1]:
        vars.update(_bind_formals(formals, it.eval(expr, env)))
    return _body_tail(it, x[2:], Env(env, vars))


def sf_let_star_values(it, x, env):
    new = env
    for formals, expr in x[1]:
        new = Env(new, _bind_formals(formals, it.eval(expr, new)))
    return _body_tail(it, x[2:], Env(new))


def sf_let_optionals(it, x, env):
    args = list(it.eval(x[1], env))
    new = env
    for spec in x[2]:
        var, default = (spec[0], spec[1]) if isinstance(spec, list) else (spec, None)
        new = Env(new, {var: args.pop(0) if args else it.eval(default, new)})
    return _body_tail(it, x[3:], Env(new))


def sf_let1(it, x, env):
    return _body_tail(it, x[3:], Env(env, {x[1]: it.eval(x[2], env)}))


def _update(fn):
    def sf(it, x, env):
        delta = it.eval(x[2], env) if len(x) > 2 else 1
        if isinstance(x[1], list):
            raise SchemeError("inc!/dec! on a generalised place is not supported")
        val = fn(lookup(x[1], env), delta)
        set_var(x[1], env, val)
        return val
    return sf


def sf_push(it, x, env):
    set_var(x[1], env, Pair(it.eval(x[2], env), lookup(x[1], env)))
    return None


def sf_pop(it, x, env):
    p = lookup(x[1], env)
    set_var(x[1], env, p.cdr)
    return p.car


def sf_define_syntax(it, x, env):
    rules_form = x[2]
    assert symname(rules_form[0]) == "syntax-rules", "only syntax-rules macros are supported"
    rules = [(r[0], r[1]) for r in rules_form[2:]]
    for pattern, template in rules:
        if any(symname(s) == "..." for s in _flatten(pattern)):
            raise SchemeError("syntax-rules ellipsis is not supported")
    module_of(env).vars[symname(x[1])] = Macro(rules, module_of(env))
    return None


def _flatten(x):
    if isinstance(x, list):
        for y in x:
            yield from _flatten(y)
    else:
        yield x


def sf_define_module(it, x, env):
    name = x[1]
    m = it.modules.get(name)
    if m is None:
        m = it.modules[name] = Module(name, it)
    for clause in x[2:]:
        it.eval(clause, m)
    return None


def sf_use(it, x, env):
    prefix = None
    rest = x[2:]
    while rest:
        if rest[0] == ":prefix":
            prefix = str(rest[1])
        rest = rest[2:]
    it.use(module_of(env), str(x[1]), prefix)
    return None


def sf_export(it, x, env):
    module_of(env).exports.update(symname(s) for s in x[1:])
    return None


def sf_export_all(it, x, env):
    module_of(env).export_all = True
    return None


def sf_ignore(it, x, env):
    return None


SPECIAL = {
    "quote": sf_quote, "if": sf_if, "define": sf_define, "define-inline": sf_define, "define-constant": sf_define, "lambda": sf_lambda,
    "let": sf_let, "let*": sf_let_star, "letrec": sf_letrec, "letrec*": sf_letrec, "set!": sf_set, "begin": sf_begin, "cond": sf_cond,
    "case": sf_case, "and": sf_and, "or": sf_or, "when": sf_when, "unless": sf_unless, "dotimes": sf_dotimes, "do": sf_do,
    "receive": sf_receive, "let-values": sf_let_values, "let*-values": sf_let_star_values, "let-optionals*": sf_let_optionals, "let1": sf_let1,
    "inc!": _update(lambda a, b: n_add(a, b)), "dec!": _update(lambda a, b: n_sub(a, b)), "push!": sf_push, "pop!": sf_pop,
    "define-syntax": sf_define_syntax, "define-module": sf_define_module, "use": sf_use, "export": sf_export, "export-all": sf_export_all,
    "define-class": sf_ignore, "add-load-path": sf_ignore, "select-module": sf_ignore, "import": sf_ignore,
}
SPECIAL = {Sym(k): v for k, v in SPECIAL.items()}
