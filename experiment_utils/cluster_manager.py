"""Alias (``experiment_utils/cluster_manager.py``)."""
from stochastic_gradient_push_b200.experiment.cluster_manager import *  # noqa: F401,F403
