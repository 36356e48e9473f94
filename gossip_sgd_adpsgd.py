#!/usr/bin/env python
"""Entry point: AD-PSGD (asynchronous bilateral gossip) ResNet trainer.
See ``stochastic_gradient_push_b200/cli/gossip_sgd_adpsgd.py`` (flags = the
reference's ``gossip_sgd_adpsgd.py``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from stochastic_gradient_push_b200.cli.gossip_sgd_adpsgd import (  # noqa: E402,F401
    main, train, validate, parse_args, update_global_iteration_counter,
    update_bilat_learning_rate)
from stochastic_gradient_push_b200.cli.common import (  # noqa: E402,F401
    accuracy, update_state, make_dataloader, init_model)

if __name__ == '__main__':
    main()
