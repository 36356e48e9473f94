"""Alias (``experiment_utils/helpers.py``)."""
from stochastic_gradient_push_b200.experiment.helpers import *  # noqa: F401,F403
