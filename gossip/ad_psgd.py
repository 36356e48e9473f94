"""Alias of :mod:`stochastic_gradient_push_b200.parallel.ad_psgd` (reference module path ``gossip/ad_psgd.py``)."""
import sys as _sys
import stochastic_gradient_push_b200.parallel.ad_psgd as _impl
_sys.modules[__name__] = _impl
