"""Model tracing and C emission (stand-in for Symbolics.jl's build_function C target).

The reference traces user functions with Symbolics.jl at construction time and
``eval``s in-place Julia callables ``fn(out, x, u, w)``
(/root/reference/src/dynamics.jl:16-34, src/costs.jl:17-44,
src/constraints.jl:17-43).  Julia is not available in this image, so the same
job is done with sympy: the user's Python function is called on sympy symbols,
derivatives are taken symbolically, common sub-expressions are eliminated and
one C header is emitted.  The header is compiled twice from the same text:

* by nvcc for sm_100a, inlined into the engine kernels (csrc/ilqr_engine.cu);
* by gcc for the CPU oracle (oracle/ilqr_oracle.c) -- test infrastructure only.

All matrices are written column-major and flat, like the reference's Julia
``Matrix{Float64}`` caches (src/dynamics.jl:31-33).
"""
from __future__ import annotations

import hashlib
import os
from typing import Callable, Iterable, Sequence

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter

CODEGEN_VERSION = "5"


# --------------------------------------------------------------------------- tracing
def _symbols(prefix: str, count: int):
    return [sp.Symbol(f"{prefix}{i}", real=True) for i in range(count)]


def _as_list(y) -> list:
    if isinstance(y, sp.MatrixBase):
        return [sp.sympify(e) for e in y]
    if isinstance(y, np.ndarray):
        return [sp.sympify(e) for e in y.ravel().tolist()]
    if isinstance(y, (list, tuple)):
        out = []
        for e in y:
            out.extend(_as_list(e))
        return out
    return [sp.sympify(y)]


class SymVec(list):
    """What a traced user function receives for x, u, w: a list of sympy symbols
    with just enough array sugar (slicing, +, -, scalar *) to write models the way
    the reference's examples do (examples/acrobot.jl:18-88)."""

    def __getitem__(self, idx):
        r = list.__getitem__(self, idx)
        return SymVec(r) if isinstance(idx, slice) else r

    def _zip(self, other, op):
        if isinstance(other, (list, tuple, np.ndarray)):
            other = list(other)
            if len(other) != len(self):
                raise ValueError("length mismatch")
            return SymVec(op(a, b) for a, b in zip(self, other))
        return SymVec(op(a, other) for a in self)

    def __add__(self, o):
        return self._zip(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._zip(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._zip(o, lambda a, b: b - a)

    def __mul__(self, o):
        return self._zip(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._zip(o, lambda a, b: a / b)

    def __neg__(self):
        return SymVec(-a for a in self)


def dot(a, b):
    """LinearAlgebra.dot for traced vectors."""
    a, b = list(a), list(b)
    if len(a) != len(b):
        raise ValueError("dot: length mismatch")
    return sum((x * y for x, y in zip(a, b)), sp.Integer(0))


def vcat(*parts):
    """Julia's [a; b; c] for traced vectors / scalars."""
    out = SymVec()
    for p in parts:
        out.extend(_as_list(p))
    return out


def trace(f: Callable, n: int, m: int, p: int):
    """Call f(x, u[, w]) on symbols, like src/dynamics.jl:18-23."""
    x, u, w = SymVec(_symbols("x", n)), SymVec(_symbols("u", m)), SymVec(_symbols("w", p))
    y = f(x, u, w) if p > 0 else f(x, u)
    return _as_list(y), x, u, w


def jacobian(exprs: Sequence[sp.Expr], vars_: Sequence[sp.Symbol]) -> sp.Matrix:
    return sp.Matrix(len(exprs), len(vars_), lambda i, j: sp.diff(exprs[i], vars_[j]))


# --------------------------------------------------------------------------- printing
class _Printer(C99CodePrinter):
    """C printer restricted to operations that are bit-reproducible on host and
    device (see include/ilqr_model_rt.h)."""

    def _print_Float(self, expr):
        return repr(float(expr))

    def _print_Integer(self, expr):
        return f"{int(expr)}.0"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"

    def _print_Pow(self, expr):
        base, exp = expr.base, expr.exp
        if exp.is_Integer or (exp.is_Float and float(exp) == int(exp)):
            k = int(exp)
            b = self.parenthesize(base, 1000)
            if k == 0:
                return "1.0"
            body = "*".join([b] * abs(k))
            if abs(k) > 1:
                body = f"({body})"
            return body if k > 0 else f"(1.0/{body})"
        if exp == sp.Rational(1, 2) or (exp.is_Float and float(exp) == 0.5):
            return f"sqrt({self._print(base)})"
        if exp == sp.Rational(-1, 2) or (exp.is_Float and float(exp) == -0.5):
            return f"(1.0/sqrt({self._print(base)}))"
        # not bit-reproducible across host/device; allowed, but parity degrades to tolerance
        return f"pow({self._print(base)}, {self._print(exp)})"

    def _print_Add(self, expr, order=None):
        """Sums are emitted as explicit fused multiply-add chains: every term that is a product
        becomes ilqr_fma(a, b, acc).  Explicit calls (rather than compiler contraction, which is
        switched off) keep host and device bit-identical while shortening the dependent chains
        the sequential kernels are bound by."""
        terms = list(expr.as_ordered_terms())
        prods, others = [], []
        for t in terms:
            factors = list(sp.Mul.make_args(t))
            if len(factors) >= 2 and not (len(factors) == 2 and factors[0] == -1):
                prods.append(factors)
            else:
                others.append(t)
        if not prods:
            return super()._print_Add(expr, order=order)
        if others:
            acc = super()._print_Add(sp.Add(*others, evaluate=False), order=order) if len(others) > 1 else self._print(others[0])
        else:
            first = prods.pop(0)
            acc = self._print(sp.Mul(*first))
        for factors in prods:
            if factors[0] == -1 and len(factors) >= 3:
                a = "-(" + self._print(factors[1]) + ")"
                b = self._print(sp.Mul(*factors[2:]))
            else:
                a = self._print(factors[0])
                b = self._print(sp.Mul(*factors[1:]))
            acc = f"ilqr_fma({a}, {b}, {acc})"
        return acc

    # sin/cos atoms whose value is already held in a named temporary (set by _emit_body).  They are printed by
    # name here, NOT substituted into the expression, so the term order of the enclosing sums (and with it the
    # rounding of the fma chains) stays that of the traced expression.
    trig_names: dict = {}

    def _print_sin(self, expr):
        hit = self.trig_names.get(expr.args[0])
        return hit[0] if hit else f"ilqr_sin({self._print(expr.args[0])})"

    def _print_cos(self, expr):
        hit = self.trig_names.get(expr.args[0])
        return hit[1] if hit else f"ilqr_cos({self._print(expr.args[0])})"


_printer = _Printer()


TABLE_MIN_NNZ = 1024  # ... or shorter arrays with at least this many coefficient entries (a dense plant matrix)
DENSE_MAX_ROWS = 128  # dense coefficient blocks up to this many rows keep their accumulators in registers
TABLE_MIN = int(os.environ.get("ILQR_TABLE_MIN", "256"))  # output arrays at least this long are emitted as constant tables + a loop
                                                          # when every entry is affine in already-computed values (dense linear models)
TRIG_GROUP = int(os.environ.get("ILQR_TRIG_GROUP", "6"))  # independent sin/cos arguments evaluated behind one range test


def _emit_trig_group(lines: list[str], group: list[tuple[str, str, str]]):
    """group: (arg_symbol, sin_symbol, cos_symbol).  One argument: plain ilqr_sincos.  Several: one joint range
    test, then the branch-free bodies back to back (the in-order SM interleaves the independent chains); any
    argument outside the fast range sends the whole group through ilqr_sincos.  Same bits either way."""
    decl = ", ".join(f"{s}, {c}" for _, s, c in group)
    lines.append(f"    double {decl};")
    if len(group) == 1:
        a, s, c = group[0]
        lines.append(f"    ilqr_sincos({a}, &{s}, &{c});")
        return
    test = " & ".join(f"ilqr_trig_is_small({a})" for a, _, _ in group)
    lines.append(f"    if ({test}) {{")
    lines += [f"        ilqr_sincos_small({a}, &{s}, &{c});" for a, s, c in group]
    lines.append("    } else {")
    lines += [f"        ilqr_sincos({a}, &{s}, &{c});" for a, s, c in group]
    lines.append("    }")


_dense_parts: dict = {}  # output name -> row-sliced variant of its dense table block (filled by _emit_table)
_slice_parts: dict = {}  # output name -> entry-sliced variant of its table block, any shape (CTA-per-step linearisation)
_SLICE_OUT = "[i_ + (i_ / ILQR_N) * colpad_]"  # column j of an ILQR_N-row output starts at j (ILQR_N + colpad_)


def _affine_rows(exprs, known=None):
    """[(const, [(coef, term), ...]), ...] if every expression is  const + sum coef * term  with numeric coefficients and
    terms that are symbols or products of symbols, else None"""
    rows = []
    for e in exprs:
        e = sp.sympify(e)
        const, terms = 0.0, []
        for term, coef in e.as_coefficients_dict().items():
            if not coef.is_number:
                return None
            if term == 1:
                const += float(coef)
            elif isinstance(term, sp.Symbol) or (isinstance(term, sp.Mul) and all(isinstance(f, sp.Symbol) for f in term.args)):
                if known is not None and term in known:
                    term = known[term]  # the product already has a name (CSE extracted it for other rows): one term, not two
                terms.append((float(coef), term))  # a value that exists already, or a plain product of such
            else:
                return None
        rows.append((const, sorted(terms, key=lambda ct: str(ct[1]))))
    return rows


def _emit_table(name: str, rows) -> tuple[list[str], list]:
    """C text computing name[i] = const_i + sum_k coef_ik * symbol_k from CSR tables (one fma chain per entry, terms in a
    fixed order), and the symbols whose values must exist before it.  Thousands of generated statements become three
    constant arrays and a loop: the 64 x 64 plant of BASELINE config 4 compiles in seconds instead of half an hour."""
    syms = sorted({t for _, terms in rows for _, t in terms}, key=str)
    col_of = {t: i for i, t in enumerate(syms)}
    printed_t = [_printer.doprint(t) for t in syms]  # the values the entries are affine in
    ptr, col, val = [0], [], []
    for _, terms in rows:
        for c, t in terms:
            col.append(col_of[t]); val.append(c)
        ptr.append(len(col))
    n = len(rows)
    lines = ["    {"]
    if syms and n <= DENSE_MAX_ROWS and len(col) * 2 >= n * len(syms):
        # dense block (a plant matrix): coefficient table [term][row], TERM loop outside, the rows' accumulators unrolled
        # inside -- n independent fma chains in flight instead of one chain per row walked entry by entry (each row still
        # adds its terms in ascending term order, entries that are absent contribute an exact +0)
        tab = [[0.0] * n for _ in syms]
        for i, (_, terms) in enumerate(rows):
            for c, t in terms:
                tab[col_of[t]][i] = c
        flat = ", ".join(repr(float(v)) for r in tab for v in r)
        lines.append(f"        static const double {name}_c0[{n}] = {{{', '.join(repr(float(c)) for c, _ in rows)}}};")
        lines.append(f"        static const double {name}_tab[{len(syms) * n}] = {{{flat}}};")
        lines.append(f"        const double {name}_t[{len(syms)}] = {{{', '.join(_printer.doprint(t) for t in syms)}}};")
        lines.append(f"        double acc_[{n}];")
        lines.append("#ifdef __CUDA_ARCH__\n#pragma unroll\n#endif")
        lines.append(f"        for (int i_ = 0; i_ < {n}; ++i_) acc_[i_] = {name}_c0[i_];")
        lines.append("#ifdef __CUDA_ARCH__\n#pragma unroll 1\n#endif")
        lines.append(f"        for (int k_ = 0; k_ < {len(syms)}; ++k_) {{")
        lines.append(f"            const double t_k = {name}_t[k_];")
        lines.append("#ifdef __CUDA_ARCH__\n#pragma unroll\n#endif")
        lines.append(f"            for (int i_ = 0; i_ < {n}; ++i_) acc_[i_] = ilqr_fma({name}_tab[k_ * {n} + i_], t_k, acc_[i_]);")
        lines.append("        }")
        lines.append("#ifdef __CUDA_ARCH__\n#pragma unroll\n#endif")
        lines.append(f"        for (int i_ = 0; i_ < {n}; ++i_) {name}[i_] = acc_[i_];")
        lines.append("    }")
        # the same block for a SLICE of the rows (i0_, i0_ + step_, ...): what one lane of a warp-per-problem kernel computes;
        # every row is the same fma chain as above (accumulator = constant, terms in ascending order), two rows interleaved
        part = ["    {",
                f"        static const double {name}_c0[{n}] = {{{', '.join(repr(float(c)) for c, _ in rows)}}};",
                f"        static const double {name}_tab[{len(syms) * n}] = {{{flat}}};",
                f"        const double {name}_t[{len(syms)}] = {{{', '.join(_printer.doprint(t) for t in syms)}}};",
                f"        for (int i_ = i0_; i_ < {n}; i_ += 2 * step_) {{",
                f"            const int j_ = i_ + step_ < {n} ? i_ + step_ : i_;",
                f"            double a0_ = {name}_c0[i_], a1_ = {name}_c0[j_];",
                f"            for (int k_ = 0; k_ < {len(syms)}; ++k_) {{",
                f"                const double t_k = {name}_t[k_];",
                f"                a0_ = ilqr_fma({name}_tab[k_ * {n} + i_], t_k, a0_);",
                f"                a1_ = ilqr_fma({name}_tab[k_ * {n} + j_], t_k, a1_);",
                "            }",
                f"            {name}[i_] = a0_;",
                f"            {name}[j_] = a1_;",
                "        }",
                "    }"]
        _dense_parts[name] = part
        _slice_parts[name] = [printed_t] + part[:3] + [
            f"        for (int i_ = i0_; i_ < {n}; i_ += step_) {{",
            f"            double a0_ = {name}_c0[i_];",
            f"            for (int k_ = 0; k_ < {len(syms)}; ++k_) a0_ = ilqr_fma({name}_tab[k_ * {n} + i_], {name}_t[k_], a0_);",
            f"            {name}{_SLICE_OUT} = a0_;",
            "        }",
            "    }"]
        return lines, syms
    def arr(ctype, ident, values, fmt):
        body = ", ".join(fmt(v) for v in values) if values else fmt(0)
        lines.append(f"        static const {ctype} {ident}[{max(len(values), 1)}] = {{{body}}};")
    arr("double", f"{name}_c0", [c for c, _ in rows], lambda v: repr(float(v)))
    if col:
        arr("int", f"{name}_ptr", ptr, lambda v: str(int(v)))
        arr("short", f"{name}_col", col, lambda v: str(int(v)))
        arr("double", f"{name}_val", val, lambda v: repr(float(v)))
        lines.append(f"        const double {name}_t[{len(syms)}] = {{{', '.join(_printer.doprint(t) for t in syms)}}};")
        lines.append(f"        for (int i_ = 0; i_ < {n}; ++i_) {{")
        lines.append(f"            double acc_ = {name}_c0[i_];")
        lines.append(f"            for (int k_ = {name}_ptr[i_]; k_ < {name}_ptr[i_ + 1]; ++k_) acc_ = ilqr_fma({name}_val[k_], {name}_t[{name}_col[k_]], acc_);")
        lines.append(f"            {name}[i_] = acc_;")
        lines.append("        }")
    else:
        lines.append(f"        for (int i_ = 0; i_ < {n}; ++i_) {name}[i_] = {name}_c0[i_];")
    lines.append("    }")
    # the same entries for a slice i0_, i0_ + step_, ... (one thread's share when a CTA evaluates the function together)
    _slice_parts[name] = [printed_t] + [ln.replace(f"for (int i_ = 0; i_ < {n}; ++i_)", f"for (int i_ = i0_; i_ < {n}; i_ += step_)")
                                        .replace(f"{name}[i_] =", f"{name}{_SLICE_OUT} =")
                                        for ln in lines if f"const double {name}_t[" not in ln]
    if col and all(len(terms) == 1 for _, terms in rows):
        # every entry is c0 + val * t[col] (a plant matrix scaled column by column, say): no row pointers and no inner loop
        # in the sliced variant, so that the entries of one caller overlap (same single fma per entry as the CSR walk)
        flat = [ln for ln in _slice_parts[name][1:] if f"{name}_ptr[" not in ln and "acc_" not in ln and f"for (int i_ = i0_" not in ln
                and ln.strip() not in ("}", "{")]
        _slice_parts[name] = [printed_t, "    {"] + flat + [
            "#ifdef __CUDA_ARCH__\n#pragma unroll 4\n#endif",
            f"        for (int i_ = i0_; i_ < {n}; i_ += step_)",
            f"            {name}{_SLICE_OUT} = ilqr_fma({name}_val[i_], {name}_t[{name}_col[i_]], {name}_c0[i_]);",
            "    }"]
    return lines, syms


def _emit_body(outputs: list[tuple[str, list[sp.Expr]]], tmp_prefix: str) -> list[str]:
    """CSE over all outputs of one function group, then print statements.  Every sin/cos becomes a symbol fed by
    a sincos evaluation; evaluations whose arguments do not depend on one another are hoisted into groups
    (level by level) in front of the statements that consume them."""
    flat = [e for _, es in outputs for e in es]
    if not flat:
        return []
    syms = sp.numbered_symbols(tmp_prefix)
    repl, red = sp.cse(flat, symbols=syms, order="canonical")

    stmts: list[tuple[str, sp.Expr]] = [(str(sym), e) for sym, e in repl]
    k = 0
    tables: list[tuple[str, list]] = []  # long affine output arrays: emitted as constant tables + a loop (_emit_table)
    for name, es in outputs:
        rows = _affine_rows(red[k:k + len(es)], known={e: sym for sym, e in repl}) if len(es) >= 16 else None
        if rows is not None and len(es) < TABLE_MIN and sum(len(t) for _, t in rows) < TABLE_MIN_NNZ:
            rows = None  # short and sparse: straight-line statements are fine
        if rows is not None:
            tables.append((name, rows))
        else:
            for i, _ in enumerate(es):
                stmts.append((f"{name}[{i}]", red[k + i]))
        k += len(es)
    index_of = {sym: i for i, (sym, _) in enumerate(repl)}

    trig_args = sorted({a.args[0] for _, e in stmts for a in e.atoms(sp.sin, sp.cos)}, key=sp.default_sort_key)
    names = {arg: (f"{tmp_prefix}a{i}", f"{tmp_prefix}s{i}", f"{tmp_prefix}c{i}") for i, arg in enumerate(trig_args)}
    # arguments needing both sin and cos are substituted by symbols (as generator versions <= 5 did, which fixes
    # the term order of the sums they appear in); lone sin / cos atoms are printed by name (see _Printer)
    kinds: dict = {}
    for _, e in stmts:
        for a in e.atoms(sp.sin, sp.cos):
            kinds.setdefault(a.args[0], set()).add(type(a))
    subs = {}
    by_name = {}
    for arg, (_, s_, c_) in names.items():
        if len(kinds[arg]) == 2:
            subs[sp.sin(arg)] = sp.Symbol(s_)
            subs[sp.cos(arg)] = sp.Symbol(c_)
        else:
            by_name[arg] = (s_, c_)
    _printer.trig_names = by_name

    # level = number of sincos rounds that must precede an expression
    stmt_level: dict[int, int] = {}
    trig_level: dict = {}

    def level_of_expr(e, own_trig: bool) -> int:
        lv = 0
        for fs in e.free_symbols:
            if fs in index_of:
                lv = max(lv, level_of_stmt(index_of[fs]))
        if own_trig:
            for a in e.atoms(sp.sin, sp.cos):
                lv = max(lv, level_of_trig(a.args[0]) + 1)
        return lv

    def level_of_stmt(i: int) -> int:
        if i not in stmt_level:
            stmt_level[i] = level_of_expr(stmts[i][1], True)
        return stmt_level[i]

    def level_of_trig(arg) -> int:
        if arg not in trig_level:
            trig_level[arg] = level_of_expr(arg, True)
        return trig_level[arg]

    lines: list[str] = []
    done: set[int] = set()

    def emit_stmt(i: int):
        if i in done:
            return
        done.add(i)
        lhs, e = stmts[i]
        for fs in sorted((index_of[f] for f in e.free_symbols if f in index_of)):
            emit_stmt(fs)
        decl = "" if "[" in lhs else "const double "
        lines.append(f"    {decl}{lhs} = {_printer.doprint(e.xreplace(subs))};")

    def emit_deps_of(e):
        for fs in sorted((index_of[f] for f in e.free_symbols if f in index_of)):
            emit_stmt(fs)

    for lv in range(max([level_of_trig(a) for a in trig_args], default=-1) + 1):
        batch = [a for a in trig_args if level_of_trig(a) == lv]
        for g0 in range(0, len(batch), max(TRIG_GROUP, 1)):
            group = batch[g0:g0 + max(TRIG_GROUP, 1)]
            for arg in group:
                emit_deps_of(arg)
                lines.append(f"    const double {names[arg][0]} = {_printer.doprint(arg.xreplace(subs))};")
            _emit_trig_group(lines, [names[arg] for arg in group])
    for i in range(len(stmts)):
        emit_stmt(i)
    for name, rows in tables:  # every temporary has been emitted by now
        tl, _ = _emit_table(name, rows)
        lines.extend(tl)
    _printer.trig_names = {}
    return lines


def _emit_function(name: str, outputs: list[tuple[str, list[sp.Expr]]], with_part: bool = False, with_slices: bool = False) -> str:
    args = ", ".join(f"double* __restrict__ {o}" for o, _ in outputs)
    _dense_parts.clear()
    _slice_parts.clear()
    body = _emit_body(outputs, "t_")
    # big straight-line functions (dense models) are compiled once and called, not inlined at every use
    qual = "ILQR_HD_NOINLINE" if len(body) > 400 or any("static const" in ln for ln in body) else "ILQR_HD"
    sig = (f"{qual} void {name}({args}, const double* __restrict__ x, "
           f"const double* __restrict__ u, const double* __restrict__ w)")
    used = "\n".join(body)
    voids = "".join(f" (void){v};" for v in ("x", "u", "w"))
    text = f"{sig} {{\n   {voids}\n{used}\n}}\n"
    if with_part and len(outputs) == 1 and outputs[0][0] in _dense_parts:
        # {name}_part: rows i0_, i0_ + step_, ... of the single dense output (warp-per-problem kernels: one slice per lane)
        oname = outputs[0][0]
        full = _emit_table_free_body(body, oname)
        psig = (f"ILQR_HD_NOINLINE void {name}_part({args}, const double* __restrict__ x, const double* __restrict__ u, "
                f"const double* __restrict__ w, int i0_, int step_)")
        text += f"#define ILQR_HAVE_{name.upper()}_PART 1\n{psig} {{\n   {voids}\n" + "\n".join(full + _dense_parts[oname]) + "\n}\n"
    if with_slices and all(o in _slice_parts for o, _ in outputs):
        # {name}_part: entries i0_, i0_ + step_, ... of EVERY output (all of them table blocks), written with padded columns --
        # what one thread computes when a whole CTA evaluates the function for one (problem, time step).  The values the
        # entries are affine in are computed once per call site by {name}_part_t into a caller-provided array (shared
        # memory) instead of once per thread.  Each entry is the same fma chain as in {name}.
        full = _emit_table_free_body(body, outputs[0][0])
        if not full:  # (entries built from CSE temporaries would need those exported too: not emitted then)
            tl, pl, off = [], [], 0
            for o, _ in outputs:
                printed_t, lines_o = _slice_parts[o][0], _slice_parts[o][1:]
                tl += [(off + j, e) for j, e in enumerate(printed_t)]
                pl += [lines_o[0], f"        const double* {o}_t = t_ + {off}; (void){o}_t;"] + lines_o[1:]
                off += len(printed_t)
            tsig = (f"ILQR_HD_NOINLINE void {name}_part_t(double* __restrict__ t_, const double* __restrict__ x, "
                    f"const double* __restrict__ u, const double* __restrict__ w, int i0_, int step_)")
            # caller i0_ of step_ computes the values j = i0_ (mod step_): one each when there are enough callers
            tbody = ([f"    if (step_ >= {max(off, 1)}) {{"] + [f"        if (i0_ == {j}) t_[{j}] = {e};" for j, e in tl] +
                     ["    } else if (i0_ == 0) {"] + [f"        t_[{j}] = {e};" for j, e in tl] + ["    }"])
            psig = f"ILQR_HD_NOINLINE void {name}_part({args}, const double* __restrict__ t_, int i0_, int step_, int colpad_)"
            text += (f"#define ILQR_HAVE_{name.upper()}_PART 1\n#define ILQR_{name.upper()}_PART_NT {max(off, 1)}\n"
                     f"{tsig} {{\n   {voids} (void)t_; (void)i0_; (void)step_;\n" + "\n".join(tbody) + "\n}\n"
                     f"{psig} {{\n    (void)t_;\n" + "\n".join(pl) + "\n}\n")
    return text


def _emit_table_free_body(body: list[str], oname: str) -> list[str]:
    """the statements of a function body that precede the table block of `oname` (the temporaries its terms are built from)"""
    for i, ln in enumerate(body):
        if ln.strip() == "{" and i + 1 < len(body) and f"static const double {oname}_c0[" in body[i + 1]:
            return body[:i]
    return body


def _colmajor(mat: sp.Matrix) -> list[sp.Expr]:
    r, c = mat.shape
    return [mat[i, j] for j in range(c) for i in range(r)]


def _subs_symbols(exprs, src, dst):
    mapping = dict(zip(src, dst))
    return [e.xreplace(mapping) for e in exprs]


def emit_header(name: str, dyn, cost_s, cost_T, con_s, con_T) -> str:
    """Emit the model header consumed by csrc/ilqr_engine.cu and oracle/ilqr_oracle.c.

    ``dyn`` ... ``con_T`` are the traced objects from ``api.py`` (Dynamics, Cost,
    Cost, Constraint, Constraint); the two kinds -- stage (t < T) and terminal
    (t = T, no action) -- follow SURVEY.md Q13 (src/costs.jl:63,76;
    src/constraints.jl:82)."""
    n, m, p = dyn.num_state, dyn.num_action, dyn.num_parameter
    cs, ct = con_s.num_constraint, con_T.num_constraint
    x, u, w = _symbols("x", n), _symbols("u", m), _symbols("w", p)

    def ex(obj, exprs):
        return _subs_symbols(exprs, list(obj.x) + list(obj.u) + list(obj.w), x[: len(obj.x)] + u[: len(obj.u)] + w[: len(obj.w)])

    def idx(vars_):
        return {v: i for i, v in enumerate(vars_)}

    # rename to array accesses
    arr = {}
    for pref, vs in (("x", x), ("u", u), ("w", w)):
        for i, v in enumerate(vs):
            arr[v] = sp.Symbol(f"{pref}[{i}]", real=True)

    def A(exprs):
        return [sp.sympify(e).xreplace(arr) for e in exprs]

    def raw(fname, outs, *bodies):
        """a function given as C statement bodies (explicit-derivative constructors, api.dynamics_from_c / constraint_from_c)"""
        args = ", ".join(f"double* __restrict__ {o}" for o in outs)
        text = "\n".join("    " + ln for body in bodies for ln in str(body).strip().splitlines())
        return (f"ILQR_HD void {fname}({args}, const double* __restrict__ x, const double* __restrict__ u, "
                f"const double* __restrict__ w) {{\n    (void)x; (void)u; (void)w;\n{text}\n}}\n")

    parts = []
    if getattr(dyn, "raw_c", None) is not None:
        parts.append(raw("ilqr_dyn", ["y"], dyn.raw_c.evaluate))
        parts.append(raw("ilqr_dyn_jac", ["fx", "fu"], dyn.raw_c.jacobian_state, dyn.raw_c.jacobian_action))
    else:
        parts.append(_emit_function("ilqr_dyn", [("y", A(ex(dyn, dyn.y)))], with_part=True))
        parts.append(_emit_function("ilqr_dyn_jac", [("fx", A(ex(dyn, _colmajor(dyn.fx)))), ("fu", A(ex(dyn, _colmajor(dyn.fu))))],
                                    with_slices=True))
    parts.append(_emit_function("ilqr_cost_s", [("g", A(ex(cost_s, [cost_s.g])))]))
    parts.append(_emit_function("ilqr_cost_s_grad", [
        ("gx", A(ex(cost_s, list(cost_s.gx)))), ("gu", A(ex(cost_s, list(cost_s.gu)))),
        ("gxx", A(ex(cost_s, _colmajor(cost_s.gxx)))), ("guu", A(ex(cost_s, _colmajor(cost_s.guu)))),
        ("gux", A(ex(cost_s, _colmajor(cost_s.gux))))]))
    # first-order part alone: what a step needs when the Hessians live in one accumulator per problem (HACC)
    parts.append(_emit_function("ilqr_cost_s_grad1", [
        ("gx", A(ex(cost_s, list(cost_s.gx)))), ("gu", A(ex(cost_s, list(cost_s.gu))))]))
    parts.append(_emit_function("ilqr_cost_T", [("g", A(ex(cost_T, [cost_T.g])))]))
    parts.append(_emit_function("ilqr_cost_T_grad", [
        ("gx", A(ex(cost_T, list(cost_T.gx)))), ("gxx", A(ex(cost_T, _colmajor(cost_T.gxx))))]))
    if cs > 0 and getattr(con_s, "raw_c", None) is not None:
        parts.append(raw("ilqr_con_s", ["c"], con_s.raw_c.evaluate))
        parts.append(raw("ilqr_con_s_jac", ["cx", "cu"], con_s.raw_c.jacobian_state, con_s.raw_c.jacobian_action))
    elif cs > 0:
        parts.append(_emit_function("ilqr_con_s", [("c", A(ex(con_s, con_s.c)))]))
        parts.append(_emit_function("ilqr_con_s_jac", [("cx", A(ex(con_s, _colmajor(con_s.cx)))), ("cu", A(ex(con_s, _colmajor(con_s.cu))))]))
    if ct > 0 and getattr(con_T, "raw_c", None) is not None:
        parts.append(raw("ilqr_con_T", ["c"], con_T.raw_c.evaluate))
        parts.append(raw("ilqr_con_T_jac", ["cx"], con_T.raw_c.jacobian_state))
    elif ct > 0:
        parts.append(_emit_function("ilqr_con_T", [("c", A(ex(con_T, con_T.c)))]))
        parts.append(_emit_function("ilqr_con_T_jac", [("cx", A(ex(con_T, _colmajor(con_T.cx))))]))

    def ineq_fn(fname, count, ineq: Iterable[int]):
        ineq = sorted(set(int(i) for i in ineq))
        for i in ineq:
            if not 0 <= i < count:
                raise ValueError(f"indices_inequality entry {i} out of range 0..{count - 1}")
        cases = "".join(f" case {i}:" for i in ineq)
        body = f"switch (i) {{{cases} return 1; default: return 0; }}" if ineq else "(void)i; return 0;"
        return f"ILQR_HD int {fname}(int i) {{ {body} }}\n"

    parts.append(ineq_fn("ilqr_ineq_s", cs, con_s.indices_inequality))
    parts.append(ineq_fn("ilqr_ineq_T", ct, con_T.indices_inequality))

    # Stage-cost Hessians without free symbols (quadratic costs with constant weights): the engine then keeps ONE
    # Hessian accumulator per problem instead of one per time step (src/costs.jl:70-84 adds the same constants to
    # every step's accumulator, so they all hold the same bits) -- see HACC in csrc/ilqr_kernels.cuh.
    hess = list(cost_s.gxx) + list(cost_s.guu) + list(cost_s.gux)
    hess_const = int(all(not sp.sympify(e).free_symbols for e in hess))
    # Dynamics Jacobians without free symbols (a linear time-invariant plant): the wide-model path then keeps ONE staged
    # Jacobian block for the whole batch instead of one per (problem, time step) -- see JAC_CONST in csrc/ilqr_kernels.cuh.
    jac_const = int(getattr(dyn, "raw_c", None) is None and
                    all(not sp.sympify(e).free_symbols for e in list(_colmajor(dyn.fx)) + list(_colmajor(dyn.fu))))

    body = "\n".join(parts)
    digest = hashlib.sha256((CODEGEN_VERSION + body).encode()).hexdigest()[:16]
    head = f"""/* GENERATED by iterativelqr.jl_b200/codegen.py (v{CODEGEN_VERSION}) -- do not edit.
 * model: {name}   hash: {digest}
 * Signature convention fn(out..., x, u, w), column-major outputs: the in-place callable
 * convention of /root/reference/src/dynamics.jl:36-37, src/costs.jl:51-78, src/constraints.jl:69-83. */
#ifndef ILQR_MODEL_GEN_H
#define ILQR_MODEL_GEN_H
#include "ilqr_model_rt.h"
#define ILQR_MODEL_NAME "{name}"
#define ILQR_MODEL_HASH "{digest}"
#define ILQR_N {n}
#define ILQR_M {m}
#define ILQR_P {p}
#define ILQR_CS {cs}
#define ILQR_CT {ct}
#define ILQR_HESS_CONST {hess_const}
#define ILQR_JAC_CONST {jac_const}

"""
    return head + body + "\n#endif\n"


def header_hash(text: str) -> str:
    for line in text.splitlines()[:4]:
        if "hash:" in line:
            return line.split("hash:")[1].split()[0]
    raise ValueError("not a generated model header")


def lambdify_inplace(exprs: list[sp.Expr], shape, x, u, w):
    """Host-side in-place callable fn(out, x, u, w) (numpy), the analogue of
    eval(build_function(...)[2]) at src/dynamics.jl:26-28.  Model-definition helper
    only; the solve path never calls it."""
    fn = sp.lambdify([list(x), list(u), list(w)], list(exprs), modules="math", cse=True)

    def call(out, xv, uv=(), wv=None):
        wv = () if wv is None else wv
        vals = fn(list(np.asarray(xv, dtype=float).ravel()), list(np.asarray(uv, dtype=float).ravel()),
                  list(np.asarray(wv, dtype=float).ravel()))
        flat = np.asarray(vals, dtype=float)
        if len(shape) == 2:
            out[...] = flat.reshape(shape[1], shape[0]).T  # exprs are column-major
        else:
            out[...] = flat.reshape(shape)
        return None

    return call
