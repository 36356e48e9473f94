"""Alias of :mod:`stochastic_gradient_push_b200.gossiper` (reference module path ``gossip/gossiper.py``)."""
import sys as _sys
import stochastic_gradient_push_b200.gossiper as _impl
_sys.modules[__name__] = _impl
