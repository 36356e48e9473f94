"""Drop-in alias of the reference's ``experiment_utils`` package."""
from stochastic_gradient_push_b200.experiment import (
    ClusterManager, Meter, make_logger, get_tcp_interface_name)
