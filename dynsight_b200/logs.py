"""Minimal stand-in for ``dynsight.logs.logger`` (reference ``_internal/logs.py:65-285``).

The reference's logger (console output, history, auto-recording of every Insight to a temp
directory, zip export) is OUT OF SCOPE for the hot path.  Only the
call surface the hot-path methods use is kept: ``logger.log(msg)``, ``logger.auto_recording``,
``logger.record_data(insight)``, ``logger.get()``.
"""

from __future__ import annotations

import time
from typing import Any


class Logger:
    def __init__(self, auto_recording: bool = False, verbose: bool = False) -> None:
        self.auto_recording = auto_recording
        self.verbose = verbose
        self._log: list[str] = []
        self._recorded: list[Any] = []

    def log(self, msg: str) -> None:
        entry = f"[{time.strftime('%Y-%m-%d %H:%M:%S')}] {msg}"
        self._log.append(entry)
        if self.verbose:
            print(entry)  # noqa: T201

    def record_data(self, insight: Any) -> None:
        self._recorded.append(insight)

    def get(self) -> str:
        return "\n".join(self._log)

    def clear_history(self) -> None:
        self._log.clear()
        self._recorded.clear()


logger = Logger()
