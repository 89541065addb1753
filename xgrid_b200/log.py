"""Leveled logger for the B200 backend.

Keeps the observable contract of the reference logger (xgrid/util/logging.py:22-60) that user
code and tests rely on: class-level ``stdouts`` / ``stderrs`` sinks and ``level`` threshold that
callers may replace, one method per level, and ``dead()`` which logs and then raises a plain
``Exception`` carrying the message tuple.  Colours and console classes are out of scope
(SURVEY.md §2 row 11).
"""
from __future__ import annotations

import sys
from enum import IntEnum
from typing import NoReturn


class LogLevel(IntEnum):
    info = 0
    done = 1
    warn = 2
    fail = 3
    dead = 4


def _write(sink, text: str) -> None:
    """Sinks are file-like objects or reference-style consoles with ``println``."""
    println = getattr(sink, "println", None)
    if println is not None:
        println(text)
    else:
        sink.write(text + "\n")


class Logger:
    stdouts: list = [sys.stdout]
    stderrs: list = [sys.stderr]
    level: LogLevel = LogLevel.warn

    def __init__(self, owner: object | str) -> None:
        self.name = owner if isinstance(owner, str) else type(owner).__qualname__

    def log(self, level: LogLevel, *msg: str) -> None:
        if int(level) < int(Logger.level):
            return
        head, *rest = msg or ("",)
        lines = [f"[ {level.name} | {self.name} ] {head}", *map(str, rest)]
        for sink in (Logger.stdouts if level <= LogLevel.done else Logger.stderrs):
            for line in lines:
                _write(sink, line)

    def dead(self, *msg: str) -> NoReturn:
        # xgrid/util/logging.py:58-60 -- report, then raise Exception(message tuple)
        self.log(LogLevel.dead, *msg)
        raise Exception(msg)


def _leveled(level: LogLevel):
    def emit(self, *msg: str) -> None:
        self.log(level, *msg)
    emit.__name__ = level.name
    return emit


for _lv in (LogLevel.info, LogLevel.done, LogLevel.warn, LogLevel.fail):
    setattr(Logger, _lv.name, _leveled(_lv))
