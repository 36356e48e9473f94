#!/usr/bin/env python
"""Entry point: AR-SGD / SGP / Overlap-SGP / D-PSGD ResNet trainer.
See ``stochastic_gradient_push_b200/cli/gossip_sgd.py`` (flags = the reference's
``gossip_sgd.py``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from stochastic_gradient_push_b200.cli.gossip_sgd import (  # noqa: E402,F401
    main, train, validate, parse_args, build_model_and_optimizer)
from stochastic_gradient_push_b200.cli.common import (  # noqa: E402,F401
    accuracy, update_state, update_peers_per_itr, update_learning_rate, make_dataloader, init_model)
from stochastic_gradient_push_b200 import GRAPH_TOPOLOGIES, MIXING_STRATEGIES  # noqa: E402,F401

if __name__ == '__main__':
    main()
