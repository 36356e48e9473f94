"""Alias of :mod:`stochastic_gradient_push_b200.topology.graph_manager` (reference module path ``gossip/graph_manager.py``)."""
import sys as _sys
import stochastic_gradient_push_b200.topology.graph_manager as _impl
_sys.modules[__name__] = _impl
