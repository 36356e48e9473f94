from .schedule import PeerSchedule, make_schedule, BOOK_BUILDERS
from .graph_manager import (
    Edge, GraphManager, MAX_PEERS_PER_ITR, GRAPH_TOPOLOGIES,
    DynamicDirectedExponentialGraph, NPeerDynamicDirectedExponentialGraph,
    DynamicBipartiteExponentialGraph, DynamicDirectedLinearGraph,
    DynamicBipartiteLinearGraph, RingGraph)
