from .helpers import (flatten_tensors, unflatten_tensors, group_by_dtype,
                      communicate, make_logger, is_power_of,
                      create_process_group, contiguous_span)
from .metering import Meter
from .arena import FlatArena
