"""Membrane-model right-hand side: Python source -> CUDA device function.

The reference hands numbalsoda a numba `cfunc` compiled from the module's
`rhs_numba(t, states, values, parameters)` (src/knpemidg/membrane.py:88,
examples/*/mm_*.py).  On the GPU the same function has to be device code, so
its *Python source* (straight-line arithmetic on `states[i]`,
`parameters[j]`, math/numpy scalar functions, side-effect stores into
`parameters[...]`) is translated statement by statement to C.  The emitted
function has the signature

    __device__ __forceinline__ void NAME(double t, const double* y, double* dy, double* p)

Only what the shipped models and Gotran-generated code use is supported;
anything else raises `TranslationError` (no silent fallback).
"""
from __future__ import annotations

import ast
import inspect
import textwrap

__all__ = ["TranslationError", "python_rhs_source", "translate_rhs", "model_cuda_source",
           "model_struct_source", "models_header"]


class TranslationError(Exception):
    pass


_FUNCS = {
    "exp": "exp", "log": "log", "sqrt": "sqrt", "sin": "sin", "cos": "cos", "tan": "tan",
    "tanh": "tanh", "sinh": "sinh", "cosh": "cosh", "fabs": "fabs", "abs": "fabs",
    "absolute": "fabs", "floor": "floor", "ceil": "ceil", "log10": "log10", "log2": "log2",
    "expm1": "expm1", "log1p": "log1p", "atan": "atan", "arctan": "atan", "asin": "asin",
    "acos": "acos", "fmod": "fmod", "power": "pow", "pow": "pow", "atan2": "atan2",
    "maximum": "fmax", "minimum": "fmin", "fmax": "fmax", "fmin": "fmin",
}
_CONSTS = {"pi": "3.141592653589793", "e": "2.718281828459045", "inf": "(1.0/0.0)"}


def python_rhs_source(module_or_func):
    """Source of the RHS: accepts a module (uses rhs_numba), a numba cfunc
    (`.py_func`), our lazy wrapper, or a plain function."""
    f = getattr(module_or_func, "rhs_numba", module_or_func)
    for attr in ("py_func", "_pyfunc", "__wrapped__"):
        if hasattr(f, attr):
            f = getattr(f, attr)
            break
    return textwrap.dedent(inspect.getsource(f))


class _Emitter(ast.NodeVisitor):
    def __init__(self, args):
        self.t, self.y, self.dy, self.p = args
        self.locals = []
        self.lines = []

    # -- expressions -------------------------------------------------------
    def expr(self, n):
        if isinstance(n, ast.Constant):
            if isinstance(n.value, bool):
                return "1.0" if n.value else "0.0"
            if isinstance(n.value, (int, float)):
                return repr(float(n.value))
            raise TranslationError(f"constant {n.value!r}")
        if isinstance(n, ast.Name):
            if n.id == self.t:
                return "t"
            if n.id in self.locals:
                return "v_" + n.id
            raise TranslationError(f"unknown name {n.id}")
        if isinstance(n, ast.Subscript):
            base = n.value.id if isinstance(n.value, ast.Name) else None
            arr = {self.y: "y", self.dy: "dy", self.p: "p"}.get(base)
            if arr is None:
                raise TranslationError("subscript on " + ast.dump(n.value))
            idx = n.slice
            if isinstance(idx, ast.Constant) and isinstance(idx.value, int):
                return f"{arr}[{idx.value}]"
            raise TranslationError("non-constant subscript")
        if isinstance(n, ast.UnaryOp):
            if isinstance(n.op, ast.USub):
                return f"(-{self.expr(n.operand)})"
            if isinstance(n.op, ast.UAdd):
                return self.expr(n.operand)
            if isinstance(n.op, ast.Not):
                return f"(({self.expr(n.operand)}) == 0.0 ? 1.0 : 0.0)"
        if isinstance(n, ast.BinOp):
            a, b = self.expr(n.left), self.expr(n.right)
            if isinstance(n.op, ast.Add):
                return f"({a} + {b})"
            if isinstance(n.op, ast.Sub):
                return f"({a} - {b})"
            if isinstance(n.op, ast.Mult):
                return f"({a} * {b})"
            if isinstance(n.op, ast.Div):
                return f"({a} / {b})"
            if isinstance(n.op, ast.Mod):
                return f"knp_pymod({a}, {b})"
            if isinstance(n.op, ast.Pow):
                return self.power(n.left, n.right)
        if isinstance(n, ast.Compare) and len(n.ops) == 1:
            op = {ast.Lt: "<", ast.LtE: "<=", ast.Gt: ">", ast.GtE: ">=", ast.Eq: "==",
                  ast.NotEq: "!="}.get(type(n.ops[0]))
            if op:
                return f"(({self.expr(n.left)} {op} {self.expr(n.comparators[0])}) ? 1.0 : 0.0)"
        if isinstance(n, ast.IfExp):
            return f"(({self.expr(n.test)}) != 0.0 ? {self.expr(n.body)} : {self.expr(n.orelse)})"
        if isinstance(n, ast.Attribute) and n.attr in _CONSTS:
            return _CONSTS[n.attr]
        if isinstance(n, ast.Call):
            name = n.func.attr if isinstance(n.func, ast.Attribute) else getattr(n.func, "id", None)
            if name in ("float", "float64", "float_"):
                return self.expr(n.args[0])
            if name == "mod":
                return f"knp_pymod({self.expr(n.args[0])}, {self.expr(n.args[1])})"
            if name in ("pow", "power"):
                return self.power(n.args[0], n.args[1])
            if name in _FUNCS:
                return f"{_FUNCS[name]}({', '.join(self.expr(a) for a in n.args)})"
            raise TranslationError(f"call to {name}")
        raise TranslationError("unsupported expression " + ast.dump(n))

    def power(self, base, exponent):
        b = self.expr(base)
        if isinstance(exponent, ast.Constant) and float(exponent.value).is_integer() \
                and 1 <= int(exponent.value) <= 4:
            k = int(exponent.value)
            return f"knp_ipow{k}({b})"
        return f"pow({b}, {self.expr(exponent)})"

    # -- statements --------------------------------------------------------
    def target(self, tgt):
        if isinstance(tgt, ast.Name):
            if tgt.id in (self.t, self.y, self.dy, self.p):
                raise TranslationError("assignment to an argument")
            if tgt.id not in self.locals:
                self.locals.append(tgt.id)
            return "v_" + tgt.id
        if isinstance(tgt, ast.Subscript):
            base = tgt.value.id
            if base == self.y:
                raise TranslationError("store into states")
            return self.expr(tgt)
        raise TranslationError("unsupported assignment target")

    def stmt(self, s):
        if isinstance(s, ast.Expr):
            if isinstance(s.value, ast.Constant):      # docstring
                return
            raise TranslationError("bare expression statement")
        if isinstance(s, ast.Assign):
            rhs = self.expr(s.value)
            for tgt in s.targets:
                self.lines.append(f"{self.target(tgt)} = {rhs};")
            return
        if isinstance(s, ast.AugAssign):
            op = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}[type(s.op)]
            rhs = self.expr(s.value)
            self.lines.append(f"{self.target(s.target)} {op}= {rhs};")
            return
        if isinstance(s, (ast.Pass,)):
            return
        if isinstance(s, ast.Return) and s.value is None:
            return
        if isinstance(s, ast.If):
            self.lines.append(f"if (({self.expr(s.test)}) != 0.0) {{")
            for b in s.body:
                self.stmt(b)
            if s.orelse:
                self.lines.append("} else {")
                for b in s.orelse:
                    self.stmt(b)
            self.lines.append("}")
            return
        raise TranslationError("unsupported statement " + type(s).__name__)


def translate_rhs(source, name, qualifier=""):
    """Translate the Python RHS source text to a CUDA function (`KNP_HD` expands
    to `__host__ __device__ __forceinline__`, csrc/knp_common.h)."""
    tree = ast.parse(textwrap.dedent(source))
    fn = next((n for n in tree.body if isinstance(n, ast.FunctionDef)), None)
    if fn is None:
        raise TranslationError("no function definition found")
    args = [a.arg for a in fn.args.args]
    if len(args) != 4:
        raise TranslationError("rhs must take (t, states, values, parameters)")
    em = _Emitter(args)
    for s in fn.body:
        em.stmt(s)
    decl = "".join(f"    double v_{v};\n" for v in em.locals)
    body = "".join(f"    {ln}\n" for ln in em.lines)
    return (f"KNP_HD {qualifier}void {name}(double t, const double* __restrict__ y, "
            f"double* __restrict__ dy, double* __restrict__ p)\n{{\n{decl}{body}}}\n")


def model_cuda_source(module, name=None):
    name = name or ("rhs_" + module.__name__.split(".")[-1])
    ns = len(module.init_state_values())
    npar = len(module.init_parameter_values())
    return translate_rhs(python_rhs_source(module), name), name, ns, npar


def model_struct_source(module, struct_name):
    """`struct NAME { NS, NP, static rhs }` as consumed by knp::OdeStepKernel."""
    ns = len(module.init_state_values())
    npar = len(module.init_parameter_values())
    fn = translate_rhs(python_rhs_source(module), "rhs", qualifier="static ")
    fn = "".join("  " + ln + "\n" for ln in fn.splitlines())
    return (f"struct {struct_name} {{\n  static constexpr int NS = {ns};\n"
            f"  static constexpr int NP = {npar};\n{fn}}};\n")


def models_header(modules):
    """csrc/generated/models_gen.h for an ordered list of (name, module)."""
    out = ["// generated by knpemidg.odegen.models_header - do not edit\n#pragma once\n"
           '#include "../knp_ode.h"\nnamespace knp {\n']
    for name, mod in modules:
        out.append(model_struct_source(mod, "Model_" + name))
    out.append("}  // namespace knp\n")
    out.append(f"#define KNP_NUM_MODELS {len(modules)}\n")
    out.append("#define KNP_MODEL_LIST(X) \\\n" + " \\\n".join(
        f'  X({i}, knp::Model_{name}, "{name}")' for i, (name, _) in enumerate(modules)) + "\n")
    names = ", ".join(f'"{name}"' for name, _ in modules)
    ns = ", ".join(str(len(m.init_state_values())) for _, m in modules)
    npar = ", ".join(str(len(m.init_parameter_values())) for _, m in modules)
    out.append(f"static const char* const knp_model_names[] = {{{names}}};\n")
    out.append(f"static const int knp_model_ns[] = {{{ns}}};\n")
    out.append(f"static const int knp_model_np[] = {{{npar}}};\n")
    return "".join(out)
