"""
Tracing (the reference has wall-clock meters only).

``Tracer`` records host-side spans in the Chrome trace-event format (open the JSON in
``chrome://tracing`` or https://ui.perfetto.dev) and, on a CUDA machine, mirrors every span
as an NVTX range so the same names show up in Nsight Systems next to the kernels.  It is a
process-global that defaults to a disabled instance whose ``span()`` is a shared no-op context,
so instrumented hot paths cost one attribute check when tracing is off.

    tracer = tracing.enable('trace', rank=rank)         # -> trace_r{rank}.json on dump()
    with tracing.span('forward'):
        ...
    tracing.counter('exposed_comm_ms', 1.25)
    tracer.dump()

The training CLIs switch it on with ``--trace_file PREFIX [--trace_iters N]``.
"""

from __future__ import annotations

import contextlib
import json
import os
import threading
import time
from typing import Optional


class _NullSpan(object):
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL_SPAN = _NullSpan()


class Tracer(object):

    def __init__(self, prefix: Optional[str] = None, rank: int = 0, enabled: bool = True,
                 nvtx: Optional[bool] = None, max_events: int = 200000):
        self.prefix = prefix
        self.rank = int(rank)
        self.enabled = bool(enabled)
        self.max_events = int(max_events)
        self._events = []
        self._lock = threading.Lock()
        self._t0 = time.perf_counter()
        self._nvtx = None
        if self.enabled and nvtx is not False:
            try:
                import torch
                if torch.cuda.is_available():
                    self._nvtx = torch.cuda.nvtx
            except Exception:           # tracing must never take the job down
                self._nvtx = None

    # ------------------------------------------------------------------ #
    def _now_us(self) -> float:
        return (time.perf_counter() - self._t0) * 1e6

    def _push(self, ev: dict):
        with self._lock:
            if len(self._events) < self.max_events:
                self._events.append(ev)

    def span(self, name: str, **args):
        """Context manager: one complete ('X') event, plus an NVTX range on CUDA machines."""
        if not self.enabled:
            return _NULL_SPAN
        return self._span(name, args)

    @contextlib.contextmanager
    def _span(self, name, args):
        if self._nvtx is not None:
            self._nvtx.range_push(name)
        t = self._now_us()
        try:
            yield self
        finally:
            dur = self._now_us() - t
            if self._nvtx is not None:
                self._nvtx.range_pop()
            ev = {'name': name, 'ph': 'X', 'ts': t, 'dur': dur, 'pid': self.rank,
                  'tid': threading.get_ident() % 100000}
            if args:
                ev['args'] = args
            self._push(ev)

    def instant(self, name: str, **args):
        if self.enabled:
            self._push({'name': name, 'ph': 'i', 's': 't', 'ts': self._now_us(), 'pid': self.rank,
                        'tid': threading.get_ident() % 100000, 'args': args})

    def counter(self, name: str, value: float):
        if self.enabled:
            self._push({'name': name, 'ph': 'C', 'ts': self._now_us(), 'pid': self.rank,
                        'args': {name: float(value)}})

    # ------------------------------------------------------------------ #
    @property
    def events(self):
        with self._lock:
            return list(self._events)

    def summary(self) -> dict:
        """name -> (count, total ms) over the recorded spans."""
        out = {}
        for ev in self.events:
            if ev.get('ph') == 'X':
                n, tot = out.get(ev['name'], (0, 0.0))
                out[ev['name']] = (n + 1, tot + ev['dur'] / 1e3)
        return out

    def path(self) -> Optional[str]:
        return None if self.prefix is None else '%s_r%d.json' % (self.prefix, self.rank)

    def dump(self, path: Optional[str] = None) -> Optional[str]:
        path = path or self.path()
        if path is None or not self.enabled:
            return None
        d = os.path.dirname(path)
        if d:
            os.makedirs(d, exist_ok=True)
        meta = [{'name': 'process_name', 'ph': 'M', 'pid': self.rank, 'args': {'name': 'rank %d' % self.rank}}]
        tmp = path + '.tmp'
        with open(tmp, 'w') as f:
            json.dump({'traceEvents': meta + self.events, 'displayTimeUnit': 'ms'}, f)
        os.replace(tmp, path)
        return path


_DISABLED = Tracer(enabled=False)
_current = _DISABLED


def get_tracer() -> Tracer:
    return _current


def enable(prefix: Optional[str] = None, rank: int = 0, **kw) -> Tracer:
    global _current
    _current = Tracer(prefix, rank, enabled=True, **kw)
    return _current


def disable() -> Tracer:
    """Stop recording; returns the tracer that was active (so it can still be dumped)."""
    global _current
    prev, _current = _current, _DISABLED
    return prev


def span(name: str, **args):
    return _current.span(name, **args)


def instant(name: str, **args):
    _current.instant(name, **args)


def counter(name: str, value: float):
    _current.counter(name, value)
