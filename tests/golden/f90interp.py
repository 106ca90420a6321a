"""A tiny Fortran-90 -> Python translator, just big enough to EXECUTE THE REFERENCE'S OWN SOURCE TEXT of the
tracer-advection kernels (loop nests, assignments, if/select blocks, calls) in IEEE binary64.

Used only by tests/golden/gen_from_reference.py, in the build container where /root/reference exists, to
produce the committed golden vectors that pin the C oracle (oracle/mom5adv_oracle.c) to the reference.
It is test tooling: not imported by the product, not needed on the GPU box.

Semantics preserved:
  * every real is a Python float (IEEE binary64, no FMA contraction: CPython evaluates one op at a time);
  * `*` `/` (and `+` `-`) associate left-to-right exactly as in Fortran; unary minus differs only in where
    the exact sign flip is applied;
  * max/min are Python's builtins: the FIRST argument wins ties -- the same convention as the oracle and the
    CUDA kernels (the Fortran standard leaves max(+0.,-0.) processor dependent);
  * nint = round half away from zero; integer/real mixing follows Python == Fortran for the expressions used.
Arrays are `FArray` objects with Fortran lower bounds; `a(i,j,k)` reads an element, sections `a(:,lo:hi)`
produce rebased FArrays, element/section assignment goes through `.set`.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional, Tuple

import numpy as np


class S:
    """section marker lo:hi (None = open)"""
    __slots__ = ("lo", "hi")

    def __init__(self, lo=None, hi=None):
        self.lo, self.hi = lo, hi


class FArray:
    def __init__(self, bounds=None, fill=0.0, data=None, lo=None):
        if data is not None:
            self.a = data
            self.lo = list(lo)
        else:
            self.lo = [b[0] for b in bounds]
            shape = [b[1] - b[0] + 1 for b in bounds]
            self.a = np.full(shape[::-1], fill, dtype=np.float64)
        self.rank = len(self.lo)

    def bounds(self):
        shp = self.a.shape[::-1]
        return [(l, l + n - 1) for l, n in zip(self.lo, shp)]

    def rebase(self, lows):
        return FArray(data=self.a, lo=[(l if l is not None else 1) for l in lows])

    def _key(self, idx):
        key, newlo, sect = [], [], False
        for d, ix in enumerate(idx):
            if isinstance(ix, S):
                sect = True
                lo = self.lo[d] if ix.lo is None else ix.lo
                hi = self.lo[d] + self.a.shape[self.rank - 1 - d] - 1 if ix.hi is None else ix.hi
                key.append(slice(lo - self.lo[d], hi - self.lo[d] + 1))
                newlo.append(1)
            else:
                o = ix - self.lo[d]
                if o < 0 or o >= self.a.shape[self.rank - 1 - d]:
                    raise IndexError(f"index {ix} out of bounds in dim {d + 1} (lo={self.lo[d]})")
                key.append(o)
        return tuple(key[::-1]), newlo, sect

    def __call__(self, *idx):
        if self.rank == 3 and type(idx[0]) is int and type(idx[1]) is int and type(idx[2]) is int:
            i, j, k = idx[0] - self.lo[0], idx[1] - self.lo[1], idx[2] - self.lo[2]
            if i < 0 or j < 0 or k < 0:
                raise IndexError(f"negative offset {idx}")
            return float(self.a[k, j, i])
        key, newlo, sect = self._key(idx)
        if sect:
            return FArray(data=self.a[key], lo=newlo)
        return float(self.a[key])

    def set(self, idx, val):
        key, _, sect = self._key(idx)
        if isinstance(val, FArray):
            val = val.a
        self.a[key] = val

    def fill(self, val):
        if isinstance(val, FArray):
            val = val.a
        self.a[...] = val

    # elementwise algebra on whole arrays / sections
    def _v(self, o):
        return o.a if isinstance(o, FArray) else o

    def __neg__(self): return FArray(data=-self.a, lo=[1] * self.rank)
    def __add__(self, o): return FArray(data=self.a + self._v(o), lo=[1] * self.rank)
    def __radd__(self, o): return FArray(data=self._v(o) + self.a, lo=[1] * self.rank)
    def __sub__(self, o): return FArray(data=self.a - self._v(o), lo=[1] * self.rank)
    def __rsub__(self, o): return FArray(data=self._v(o) - self.a, lo=[1] * self.rank)
    def __mul__(self, o): return FArray(data=self.a * self._v(o), lo=[1] * self.rank)
    def __rmul__(self, o): return FArray(data=self._v(o) * self.a, lo=[1] * self.rank)
    def __truediv__(self, o): return FArray(data=self.a / self._v(o), lo=[1] * self.rank)


class FList:
    """1-based array of derived-type objects: T_prog(n)"""

    def __init__(self, items):
        self.items = list(items)

    def __call__(self, n):
        if isinstance(n, S):
            return self
        return self.items[n - 1]


class Obj:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def nint(x):
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


# ------------------------------------------------------------------------------------------------
# source handling
# ------------------------------------------------------------------------------------------------
def logical_lines(lines: List[str]) -> List[str]:
    """strip comments / cpp lines, join `&` continuations."""
    out, cur = [], ""
    for raw in lines:
        ln = raw.rstrip("\n")
        if ln.lstrip().startswith("#"):
            continue
        # strip comment (no '!' inside the string literals of the fragments we run)
        q = None
        cut = len(ln)
        for p, ch in enumerate(ln):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "!":
                cut = p
                break
        ln = ln[:cut].strip()
        if not ln:
            continue
        if ln.startswith("&"):
            ln = ln[1:].lstrip()
        if ln.endswith("&"):
            cur += ln[:-1] + " "
            continue
        out.append(cur + ln)
        cur = ""
    if cur:
        out.append(cur)
    return out


_STR = re.compile(r"'[^']*'|\"[^\"]*\"")


def _protect(s):
    strs = []

    def rep(m):
        strs.append(m.group(0))
        return f"__STR{len(strs) - 1}__"

    return _STR.sub(rep, s), strs


def _restore(s, strs):
    for n, t in enumerate(strs):
        s = s.replace(f"__STR{n}__", t)
    return s


def _split_top(s: str, sep: str) -> List[str]:
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def _xform_sections(s: str) -> str:
    """rewrite `lo:hi` arguments inside every parenthesised list as S(lo,hi)."""
    out, i = "", 0
    while i < len(s):
        ch = s[i]
        if ch == "(":
            depth, j = 1, i + 1
            while depth:
                depth += {"(": 1, ")": -1}.get(s[j], 0)
                j += 1
            inner = s[i + 1:j - 1]
            args = []
            for a in _split_top(inner, ","):
                pieces = _split_top(a, ":")
                if len(pieces) == 2:
                    lo, hi = pieces[0].strip(), pieces[1].strip()
                    args.append(f"S({_xform_sections(lo) or 'None'},{_xform_sections(hi) or 'None'})")
                else:
                    args.append(_xform_sections(a))
            out += "(" + ",".join(args) + ")"
            i = j
        else:
            out += ch
            i += 1
    return out


_LOGICAL = [(r"\.not\.", " not "), (r"\.and\.", " and "), (r"\.or\.", " or "), (r"\.true\.", " True "),
            (r"\.false\.", " False "), (r"\.eq\.", "=="), (r"\.ne\.", "!="), (r"\.lt\.", "<"), (r"\.le\.", "<="),
            (r"\.gt\.", ">"), (r"\.ge\.", ">=")]


_SQ = re.compile(r"(\((?:[^()]|\((?:[^()]|\([^()]*\))*\))*\)|[\w%]+)\s*\*\*\s*2\b")


def fsq(x):
    """x**2 as every Fortran compiler evaluates it: one multiplication (CPython's float ** calls libm pow)"""
    return x * x


def xexpr(s: str) -> str:
    s = _SQ.sub(r"fsq(\1)", s)
    s = s.replace("%", ".")
    for pat, rep in _LOGICAL:
        s = re.sub(pat, rep, s, flags=re.I)
    s = s.replace("/=", "!=")
    s = re.sub(r"\bpresent\s*\(\s*(\w+)\s*\)", r"(\1 is not None)", s, flags=re.I)
    return _xform_sections(s).strip()


def _find_assign(s: str) -> int:
    depth = 0
    for p, ch in enumerate(s):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if s[p + 1:p + 2] == "=" or s[p - 1:p] in "=<>/!":
                continue
            return p
    return -1


class Translator:
    def __init__(self, array_names=()):
        self.array_names = set(array_names)  # names that are FArrays (whole-array assignment -> .fill)

    def assign(self, lhs: str, rhs: str) -> str:
        lhs, rhs = lhs.strip(), xexpr(rhs)
        if lhs.endswith(")"):
            depth, j = 0, len(lhs) - 1
            while True:
                depth += {")": 1, "(": -1}.get(lhs[j], 0)
                if depth == 0:
                    break
                j -= 1
            obj, idx = xexpr(lhs[:j]), xexpr(lhs[j:])
            return f"{obj}.set({idx[:-1]},), {rhs})" if idx != "()" else f"{obj}.fill({rhs})"
        name = xexpr(lhs)
        if name in self.array_names or "." in name and name.split(".")[-1] in self.array_names:
            return f"{name}.fill({rhs})"
        return f"{name} = {rhs}"

    def stmt(self, s: str) -> str:
        """simple (non-block) statement"""
        m = re.match(r"call\s+(\w+)\s*(\(.*\))?\s*$", s, flags=re.I)
        if m:
            return f"{m.group(1)}{xexpr(m.group(2) or '()')}"
        if re.match(r"return\b", s, flags=re.I):
            return "return __RET__"
        p = _find_assign(s)
        if p >= 0:
            return self.assign(s[:p], s[p + 1:])
        raise SyntaxError(f"cannot translate: {s}")

    def body(self, lines: List[str], indent: int = 1) -> List[str]:
        py, ind = [], indent
        sel: List[Tuple[str, bool]] = []  # select-case stack: (selector, first-case-seen)

        def emit(t):
            py.append("    " * ind + t)

        for raw in logical_lines(lines):
            s, strs = _protect(raw)
            low = s.lower().strip()
            if re.match(r"(allocate|deallocate|use |implicit|type\s*\(|real\b|integer\b|logical\b|character\b)", low):
                continue
            m = re.match(r"do\s+(\w+)\s*=\s*(.*)$", s, flags=re.I)
            if m:
                parts = _split_top(m.group(2), ",")
                lo, hi = xexpr(parts[0]), xexpr(parts[1])
                step = xexpr(parts[2]) if len(parts) > 2 else None
                emit(f"for {m.group(1)} in range({lo}, ({hi})+1{', ' + step if step else ''}):")
                ind += 1
                emit("pass")
                continue
            if re.match(r"end\s*do\b", low):
                ind -= 1
                continue
            m = re.match(r"(else\s*)?if\s*\((.*)\)\s*then\s*$", s, flags=re.I)
            if m:
                if m.group(1):
                    ind -= 1
                    emit(f"elif {xexpr(m.group(2))}:")
                else:
                    emit(f"if {xexpr(m.group(2))}:")
                ind += 1
                emit("pass")
                continue
            if re.match(r"else\s*$", low):
                ind -= 1
                emit("else:")
                ind += 1
                emit("pass")
                continue
            if re.match(r"end\s*if\b", low):
                ind -= 1
                continue
            m = re.match(r"select\s+case\s*\((.*)\)\s*$", s, flags=re.I)
            if m:
                sel.append([xexpr(m.group(1)), False])
                continue
            m = re.match(r"case\s+default\s*$", low)
            if m:
                if sel[-1][1]:
                    ind -= 1
                emit("else:" if sel[-1][1] else "if True:")
                sel[-1][1] = True
                ind += 1
                emit("pass")
                continue
            m = re.match(r"case\s*\((.*)\)\s*$", s, flags=re.I)
            if m:
                kw = "elif" if sel[-1][1] else "if"
                if sel[-1][1]:
                    ind -= 1
                emit(f"{kw} {sel[-1][0]} == {xexpr(m.group(1))}:")
                sel[-1][1] = True
                ind += 1
                emit("pass")
                continue
            if re.match(r"end\s*select\b", low):
                if sel.pop()[1]:
                    ind -= 1
                continue
            m = re.match(r"if\s*\(", s, flags=re.I)
            if m:  # one-line if
                depth, j = 0, s.index("(")
                while True:
                    depth += {"(": 1, ")": -1}.get(s[j], 0)
                    j += 1
                    if depth == 0:
                        break
                cond, rest = s[s.index("(") + 1:j - 1], s[j:].strip()
                emit(f"if {xexpr(cond)}: " + _restore(self.stmt(rest), strs))
                continue
            emit(_restore(self.stmt(s), strs))
        return py


_DECL = re.compile(r"^\s*(real|integer|logical)\s*(,\s*dimension\s*\((?P<dims>.*?)\)\s*)?(?P<attrs>(,\s*\w+(\([^)]*\))?\s*)*)::(?P<names>.*)$", re.I)


def parse_decls(lines: List[str]):
    """-> (local arrays {name: [dim strings]}, dummy arrays {name: [lower-bound strings or None]})"""
    local, dummy = {}, {}
    for ln in logical_lines(lines):
        m = _DECL.match(ln)
        if not m:
            continue
        dims, attrs = m.group("dims"), (m.group("attrs") or "").lower()
        if dims is None:   # `real, intent(in), dimension(isd:,jsd:) :: a` -- dimension after other attributes
            md = re.search(r"dimension\s*\((.*?)\)\s*(,|$)", m.group("attrs") or "", flags=re.I)
            if md:
                dims = md.group(1)
        names = [n.strip() for n in _split_top(m.group("names"), ",")]
        for nm in names:
            d = dims
            mm = re.match(r"(\w+)\s*\((.*)\)$", nm)
            if mm:
                nm, d = mm.group(1), mm.group(2)
            if d is None:
                continue
            dl = [x.strip() for x in _split_top(d, ",")]
            if "intent" in attrs:
                dummy[nm] = [(x.split(":")[0].strip() or None) if ":" in x else "1" for x in dl]
            else:
                local[nm] = dl
    return local, dummy


def translate_routine(src_lines: List[str], first: int, last: int, array_names=()) -> Tuple[str, str]:
    """Translate the function/subroutine whose text spans src_lines[first-1:last] (1-based, inclusive).
    Returns (python source, routine name)."""
    lines = src_lines[first - 1:last]
    ll = logical_lines(lines)
    m = re.match(r"\s*(function|subroutine)\s+(\w+)\s*(\((.*)\))?", ll[0], flags=re.I)
    kind, name = m.group(1).lower(), m.group(2)
    args = [a.strip() for a in (m.group(4) or "").split(",") if a.strip()]
    local, dummy = parse_decls(lines)
    tr = Translator(set(array_names) | set(local) | set(dummy))
    # body = everything between the header and the end statement
    body_lines = []
    seen_header = False
    for raw in lines:
        if not seen_header:
            if re.match(r"\s*(function|subroutine)\s+" + name, raw, flags=re.I):
                seen_header = True
            continue
        if re.match(r"\s*end\s+(function|subroutine)", raw, flags=re.I):
            break
        body_lines.append(raw)
    # join header continuation lines away
    while body_lines and logical_lines(lines)[0].count("(") and False:
        break
    py = [f"def {name}({', '.join(a + '=None' for a in args)}):"]
    for a, lows in dummy.items():
        if a in args:
            py.append(f"    {a} = {a}.rebase([{', '.join(xexpr(l) if l else 'None' for l in lows)}]) if isinstance({a}, FArray) else {a}")
    for a, dl in local.items():
        b = []
        for x in dl:
            if ":" in x:
                lo, hi = x.split(":")
                b.append(f"({xexpr(lo)}, {xexpr(hi)})")
            else:
                b.append(f"(1, {xexpr(x)})")
        py.append(f"    {a} = FArray([{', '.join(b)}])")
    ret = name if kind == "function" else "None"
    code = tr.body(body_lines, indent=1)
    py += [c.replace("__RET__", ret) for c in code]
    py.append(f"    return {ret}")
    return "\n".join(py), name


def translate_block(src_lines: List[str], first: int, last: int, array_names=()) -> str:
    """Translate a bare statement range (no header) to module-level python."""
    tr = Translator(array_names)
    return "\n".join(tr.body(src_lines[first - 1:last], indent=0))
