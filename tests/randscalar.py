"""Random scalar-only kernels: typed expression trees over two int and two
float arguments -- C integer arithmetic (truncating division, dividend-signed remainder), casts,
comparisons, short-circuit logic, conditional expressions, `** 2.0`, locals, augmented assignment,
`for` / `while` / `if` -- kept inside defined behaviour (no signed overflow, no division by zero, float -> int
casts in range).  The reference compiles them with gcc; this backend evaluates them on the host
(`lang/schedule.py::HostEval`), which is also what computes the scalar prologue of every grid kernel."""
import numpy as np


class Gen:
    def __init__(self, seed: int) -> None:
        self.r = np.random.default_rng(seed)

    def pick(self, *options):
        return options[int(self.r.integers(len(options)))]

    def ilit(self) -> str:
        v = int(self.r.integers(-9, 10))
        return f"({v})" if v < 0 else str(v)

    def flit(self) -> str:
        v = round(float(self.r.uniform(-4, 4)), 2)
        return f"({v})" if v < 0 else str(v)

    def nonzero_ilit(self) -> str:
        v = int(self.pick(-7, -3, -2, 2, 3, 5, 7))
        return f"({v})" if v < 0 else str(v)

    def iexpr(self, d: int, ivars, fvars) -> str:
        if d <= 0 or self.r.random() < 0.25:
            return self.pick(self.ilit(), *ivars)
        k = self.r.random()
        if k < 0.30:
            return f"({self.iexpr(d - 1, ivars, fvars)} {self.pick('+', '-')} {self.iexpr(d - 1, ivars, fvars)})"
        if k < 0.42:
            return f"({self.iexpr(d - 1, ivars, fvars)} * {self.ilit()})"
        if k < 0.56:
            return f"({self.iexpr(d - 1, ivars, fvars)} {self.pick('/', '%')} {self.nonzero_ilit()})"
        if k < 0.68:
            return f"cast(int, {self.fexpr(d - 1, ivars, fvars)})"
        if k < 0.80:
            return (f"({self.iexpr(d - 1, ivars, fvars)} if {self.bexpr(d - 1, ivars, fvars)} "
                    f"else {self.iexpr(d - 1, ivars, fvars)})")
        if k < 0.90:
            return f"(-{self.iexpr(d - 1, ivars, fvars)})"
        return f"({self.iexpr(d - 1, ivars, fvars)} % {self.nonzero_ilit()} + {self.ilit()})"

    def fexpr(self, d: int, ivars, fvars) -> str:
        if d <= 0 or self.r.random() < 0.25:
            return self.pick(self.flit(), *fvars)
        k = self.r.random()
        if k < 0.35:
            return f"({self.fexpr(d - 1, ivars, fvars)} {self.pick('+', '-', '*')} {self.fexpr(d - 1, ivars, fvars)})"
        if k < 0.47:
            den = self.fexpr(d - 1, ivars, fvars)
            return f"({self.fexpr(d - 1, ivars, fvars)} / ({den} * {den} + 1.5))"
        if k < 0.59:
            return f"cast(float, {self.iexpr(d - 1, ivars, fvars)})"
        if k < 0.69:
            return f"({self.fexpr(d - 1, ivars, fvars)}) ** 2.0"
        if k < 0.81:
            return (f"({self.fexpr(d - 1, ivars, fvars)} if {self.bexpr(d - 1, ivars, fvars)} "
                    f"else {self.fexpr(d - 1, ivars, fvars)})")
        if k < 0.91:
            return f"(-{self.fexpr(d - 1, ivars, fvars)})"
        return f"({self.fexpr(d - 1, ivars, fvars)} * 0.125)"

    def bexpr(self, d: int, ivars, fvars) -> str:
        k = self.r.random()
        cmp = self.pick("<", "<=", ">", ">=", "==", "!=")
        if d <= 0 or k < 0.35:
            return f"{self.iexpr(max(d - 1, 0), ivars, fvars)} {cmp} {self.iexpr(max(d - 1, 0), ivars, fvars)}"
        if k < 0.6:
            return f"{self.fexpr(d - 1, ivars, fvars)} {cmp} {self.fexpr(d - 1, ivars, fvars)}"
        if k < 0.8:
            return f"({self.bexpr(d - 1, ivars, fvars)} {self.pick('and', 'or')} {self.bexpr(d - 1, ivars, fvars)})"
        return f"(not {self.bexpr(d - 1, ivars, fvars)})"


def gen_source(seed: int) -> tuple[str, tuple]:
    g = Gen(seed)
    ret_float = bool(g.r.random() < 0.5)
    ivars, fvars = ["a", "b"], ["x", "y"]
    lines = ["IMPORT_LINE", "from typing import cast", "", "@xgrid.kernel()",
             f"def k(a: int, b: int, x: float, y: float) -> {'float' if ret_float else 'int'}:"]
    for n in range(int(g.r.integers(1, 4))):
        if g.r.random() < 0.5:
            lines.append(f"    i{n} = {g.iexpr(3, ivars, fvars)}")
            ivars = ivars + [f"i{n}"]
        else:
            lines.append(f"    f{n} = {g.fexpr(3, ivars, fvars)}")
            fvars = fvars + [f"f{n}"]
    lines.append("    acc = 0")
    lines.append("    facc = 0.0")
    shape = g.r.random()
    if shape < 0.35:
        lines.append(f"    for t in range({g.pick(0, 1)}, {g.pick(3, 5, 6)}, {g.pick(1, 2)}):")
        lines.append(f"        if {g.bexpr(1, ivars + ['t'], fvars)}:")
        lines.append(f"            acc += {g.iexpr(2, ivars + ['t'], fvars)} % 11")
        lines.append("        else:")
        lines.append(f"            facc = facc * 0.5 + {g.fexpr(2, ivars + ['t'], fvars)}")
    elif shape < 0.6:
        lines.append("    n = 0")
        lines.append(f"    while n < {g.pick(2, 4, 7)}:")
        lines.append("        n += 1")
        lines.append(f"        if {g.bexpr(1, ivars + ['n'], fvars)}:")
        lines.append("            continue")
        lines.append(f"        acc -= {g.iexpr(2, ivars + ['n'], fvars)} / {g.nonzero_ilit()}")
        lines.append(f"        if acc > 40:")
        lines.append("            break")
    else:
        lines.append(f"    if {g.bexpr(2, ivars, fvars)}:")
        lines.append(f"        acc = {g.iexpr(2, ivars, fvars)}")
        lines.append("    else:")
        lines.append(f"        facc = {g.fexpr(2, ivars, fvars)}")
    if ret_float:
        lines.append(f"    return {g.fexpr(2, ivars + ['acc'], fvars + ['facc'])} + facc + cast(float, acc)")
    else:
        lines.append(f"    return {g.iexpr(2, ivars + ['acc'], fvars)} + acc % 1000 + cast(int, facc)")
    args = (int(g.r.integers(-50, 51)), int(g.r.integers(-50, 51)),
            round(float(g.r.uniform(-3, 3)), 3), round(float(g.r.uniform(-3, 3)), 3))
    return "\n".join(lines) + "\n", args
