"""Alias of :mod:`stochastic_gradient_push_b200.mixing_manager` (reference module path ``gossip/mixing_manager.py``)."""
import sys as _sys
import stochastic_gradient_push_b200.mixing_manager as _impl
_sys.modules[__name__] = _impl
