"""Alias of the metering module (``gossip/utils/metering.py``)."""
from stochastic_gradient_push_b200.utils.metering import Meter  # noqa: F401
