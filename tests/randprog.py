"""Random xgrid programs for differential testing (CUDA path vs oracle/interp.py).

Programs stay inside the part of the DSL whose result is well defined in the reference:
a statement on a mask value != 0 never reads level 0 of the grid it stores (the reference
races there), coefficients are small so values stay finite, divisors are non-zero scalars."""
import importlib.util
import os

import numpy as np


def _tap(rng, grids, ndim, forbid_level0_of=None, max_level=2, wide=True):
    g = grids[rng.integers(len(grids))]
    offs = []
    for a in range(ndim):
        lim = 2 if (wide and (a == ndim - 1 or a == 0) and rng.random() < 0.3) else 1
        offs.append(int(rng.integers(-lim, lim + 1)))
    r = rng.random()
    if r < 0.55:
        lvl = ""                          # default: one level back
    elif r < 0.75 and g != forbid_level0_of:
        lvl = "[0]"
    elif r < 0.9 or max_level < 2:
        lvl = "[-1]"
    else:
        lvl = "[2]" if rng.random() < 0.5 else "[-2]"
    return f"{g}[{', '.join(map(str, offs))}]{lvl}"


def _expr(rng, grids, ndim, forbid=None, max_level=2):
    n = int(rng.integers(2, 6))
    terms = []
    for _ in range(n):
        t = _tap(rng, grids, ndim, forbid, max_level)
        kind = rng.random()
        coef = round(float(rng.uniform(0.05, 0.3)), 3)
        if kind < 0.5:
            terms.append(f"{coef} * {t}")
        elif kind < 0.65:
            terms.append(f"a * {t}")
        elif kind < 0.75:
            terms.append(f"({t} - {_tap(rng, grids, ndim, forbid, max_level)}) * c")
        elif kind < 0.85:
            terms.append(f"{coef} * {t} / b")
        elif kind < 0.93:
            terms.append(f"{coef} * ({t}) ** 2.0")
        else:
            terms.append(f"({t} if a > 0.1 else {coef})")
    out = terms[0]
    for t in terms[1:]:
        out += (" + " if rng.random() < 0.7 else " - ") + t
    return out


def gen_source(seed: int, ndim: int, ngrids: int, single_1d: bool = False) -> str:
    rng = np.random.default_rng(seed)
    grids = [f"g{i}" for i in range(ngrids)]
    max_level = 1 if single_1d else 2
    lines = ["import xgrid_b200 as xgrid", f"G = xgrid.grid[float, {ndim}]", "", "@xgrid.kernel()",
             "def prog(" + ", ".join(f"{g}: G" for g in grids) + ", a: float, b: float) -> None:",
             "    c = a * b + 0.25"]
    zero = ", ".join(["0"] * ndim)
    nstmt = 2 if single_1d else int(rng.integers(2, 6))
    in_loop = False
    for s in range(nstmt):
        ind = "        " if in_loop else "    "
        tgt = grids[0] if single_1d else grids[rng.integers(len(grids))]
        r = rng.random()
        if single_1d:
            if s == 0:
                taps = [f"{tgt}[{int(rng.integers(-2, 3))}]" for _ in range(3)]
                lines.append(f"{ind}{tgt}[{zero}] = 0.3 * {taps[0]} + 0.25 * {taps[1]} + c * {taps[2]} / b")
            else:
                lines.append(f"{ind}with xgrid.boundary(1):")
                lines.append(f"{ind}    {tgt}[{zero}] = 0.5 + a")
            continue
        if r < 0.15 and not in_loop and s < nstmt - 1:
            lines.append(f"    for _ in range(0, {int(rng.integers(2, 4))}):")
            in_loop = True
            ind = "        "
        r2 = rng.random()
        if r2 < 0.12:
            # scalar control flow decides which sweep runs (host side)
            lines.append(f"{ind}if a * b > {round(float(rng.uniform(0.2, 0.8)), 2)}:")
            lines.append(f"{ind}    {tgt}[{zero}] = {_expr(rng, grids, ndim, None, max_level)}")
            lines.append(f"{ind}else:")
            lines.append(f"{ind}    {tgt}[{zero}] = {_expr(rng, grids, ndim, None, max_level)}")
        elif r2 < 0.6:
            lines.append(f"{ind}{tgt}[{zero}] = {_expr(rng, grids, ndim, None, max_level)}")
        else:
            k = int(rng.integers(1, 4))
            lines.append(f"{ind}with xgrid.boundary({k}):")
            if rng.random() < 0.5:
                lines.append(f"{ind}    {tgt}[{zero}] = {round(float(rng.uniform(-1, 1)), 2)}")
            else:
                lines.append(f"{ind}    {tgt}[{zero}] = {_expr(rng, grids, ndim, tgt, max_level)}")
        if in_loop and rng.random() < 0.5:
            in_loop = False
    return "\n".join(lines) + "\n"


def load_program(source: str, directory: str, name: str):
    path = os.path.join(directory, name + ".py")
    with open(path, "w") as f:
        f.write(source)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.prog


def gen_inputs(seed: int, shape, ngrids: int):
    rng = np.random.default_rng(seed + 1000)
    ics, masks = [], []
    for _ in range(ngrids):
        ics.append(rng.uniform(-1.0, 1.0, shape))
        m = np.zeros(shape, np.int32)
        if rng.random() < 0.8:
            for ax in range(len(shape)):
                if rng.random() < 0.7:
                    sl = [slice(None)] * len(shape)
                    sl[ax] = 0
                    m[tuple(sl)] = int(rng.integers(1, 4))
                    sl[ax] = -1
                    m[tuple(sl)] = int(rng.integers(1, 4))
        sprinkle = rng.random(shape)
        m[sprinkle < 0.01] = rng.integers(1, 4)
        m[sprinkle > 0.995] = 7                     # matches no statement
        masks.append(m)
    return ics, masks


def guard_array_ends(masks, shape, value: int = 7):
    """The reference addresses taps linearly and reads whatever lies outside the array (undefined);
    the backend reads ghost zeros there.  For comparisons AGAINST THE REFERENCE give every cell whose
    taps (offsets up to +-2 per axis) could leave the array a mask value no statement matches, so that
    no statement is evaluated there.  Row-wrapped reads inside the array stay."""
    margin = 2
    for extent in shape[1:]:
        margin = margin * extent + 2
    for m in masks:
        flat = m.reshape(-1)
        flat[:margin] = value
        flat[flat.size - margin:] = value
    return masks
