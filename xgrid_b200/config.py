"""Process-wide configuration (``xgrid.init``).

Accepts every keyword the reference accepts (xgrid/util/init.py:48) so that
existing scripts run unchanged.  ``cc`` / ``parallel`` / ``opt_level`` no
longer select a host compiler: the backend always emits CUDA C for sm_100a.
New keyword-only extras (all optional):

* ``validate``   -- build device code with ``--fmad=false`` so fp results are
                    bit-comparable with the reference's SSE2 (no-FMA) code
                    (SURVEY.md F7).  Default True: parity first.
* ``device``     -- CUDA ordinal; default = LOCAL_RANK or 0.
* ``distributed``-- slab-decompose grids over the ranks of the current
                    ``torch.distributed`` process group (one process per GPU).
* ``graphs``     -- replay steady-state kernel calls from captured CUDA graphs.
* ``temporal``   -- defer runs of identical calls of a 1-D kernel and execute them T time
                    steps per launch from shared memory (temporal blocking); results are
                    bit-identical to step-at-a-time execution.
"""
from __future__ import annotations

import os
import sys
from dataclasses import asdict, dataclass, field
from typing import Literal

from .log import Logger

_log = Logger("xgrid")


@dataclass
class Configuration:
    parallel: bool = True
    cc: list = field(default_factory=lambda: ["gcc", "clang"])
    cacheroot: str = ".xgrid"
    comment: bool = False
    overstep: Literal["none", "limit", "wrap"] = "none"
    opt_level: int = 2
    precision: Literal["float", "double"] = "float"
    # --- B200 backend extras ---
    validate: bool = True
    device: int = 0
    distributed: bool = False
    graphs: bool = True
    temporal: bool = True

    def __repr__(self) -> str:
        return repr(asdict(self))

    @property
    def fsize(self) -> int:
        # xgrid/util/init.py:34-36
        return 4 if self.precision == "float" else 8

    @property
    def nvrtc_flags(self) -> list[str]:
        flags = ["--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo",
                 "--prec-div=true", "--prec-sqrt=true", "--ftz=false"]
        flags.append("--fmad=false" if self.validate else "--fmad=true")
        return flags


_config: Configuration | None = None
_epoch = 0  # bumped by every init(); JIT artefacts are keyed on it


def get_config() -> Configuration:
    if _config is None:
        _log.dead("Please call init first to initialize")
    return _config


def config_epoch() -> int:
    return _epoch


def init(*, parallel: bool = True, cc: list[str] = ["gcc", "clang"], cacheroot: str = ".xgrid",
         comment: bool = False, overstep: Literal["none", "limit", "wrap"] = "none",
         opt_level: Literal[0, 1, 2, 3] = 2, precision: Literal["float", "double"] = "float",
         validate: bool = True, device: int | None = None, distributed: bool = False,
         graphs: bool = True, temporal: bool = True) -> None:
    global _config, _epoch
    if sys.version_info < (3, 10):
        _log.fail(f"Minimum Python 3.10 is required, current version is {sys.version_info}")
    if overstep not in ("none", "limit", "wrap"):
        _log.dead(f"Invalid overstep mode '{overstep}'")
    if precision not in ("float", "double"):
        _log.dead(f"Invalid precision '{precision}'")
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    _config = Configuration(parallel, list(cc), cacheroot, comment, overstep, opt_level, precision,
                            validate, device, distributed, graphs, temporal)
    _epoch += 1
    _log.info(f"initialized with configuration: {_config}")
