"""Alias of :mod:`stochastic_gradient_push_b200.parallel.distributed` (reference module path ``gossip/distributed.py``)."""
import sys as _sys
import stochastic_gradient_push_b200.parallel.distributed as _impl
_sys.modules[__name__] = _impl
