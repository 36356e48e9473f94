from stochastic_gradient_push_b200.utils.helpers import (
    flatten_tensors, unflatten_tensors, group_by_dtype, communicate, make_logger,
    is_power_of, create_process_group)
from stochastic_gradient_push_b200.utils import metering  # noqa: F401
from stochastic_gradient_push_b200.utils.metering import Meter  # noqa: F401
