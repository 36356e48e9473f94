"""Alias (``gossip/utils/helpers.py``)."""
from stochastic_gradient_push_b200.utils.helpers import *  # noqa: F401,F403
from stochastic_gradient_push_b200.utils.helpers import (flatten_tensors, unflatten_tensors, group_by_dtype, communicate, make_logger, is_power_of, create_process_group)  # noqa: F401
