"""Leveled logger for the B200 backend.

Mirrors the observable behaviour of the reference logger
(xgrid/util/logging.py:22-60): class-level ``stdouts`` / ``stderrs`` /
``level`` that callers may replace, and ``dead()`` which logs and then raises
a plain ``Exception`` carrying the message tuple.  Cosmetics (colours) are
intentionally minimal; SURVEY.md §2 row 11 marks them out of scope.
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


class Logger:
    stdouts: list = [sys.stdout]
    stderrs: list = [sys.stderr]
    level: LogLevel = LogLevel.warn

    def __init__(self, owner: object | str) -> None:
        self.name = owner if isinstance(owner, str) else type(owner).__qualname__

    def log(self, level: LogLevel, *msg: str) -> None:
        if int(level) < int(Logger.level):
            return
        sinks = Logger.stdouts if level <= LogLevel.done else Logger.stderrs
        for sink in sinks:
            for n, line in enumerate(msg):
                text = f"[ {level.name} | {self.name} ] {line}" if n == 0 else str(line)
                write = getattr(sink, "println", None)
                if write is not None:
                    write(text)
                else:
                    sink.write(text + "\n")

    def info(self, *msg: str) -> None:
        self.log(LogLevel.info, *msg)

    def done(self, *msg: str) -> None:
        self.log(LogLevel.done, *msg)

    def warn(self, *msg: str) -> None:
        self.log(LogLevel.warn, *msg)

    def fail(self, *msg: str) -> None:
        self.log(LogLevel.fail, *msg)

    def dead(self, *msg: str) -> NoReturn:
        # xgrid/util/logging.py:58-60 -- log, then raise Exception(msg tuple)
        self.log(LogLevel.dead, *msg)
        raise Exception(msg)
