from .cluster_manager import ClusterManager
from .helpers import make_logger, get_tcp_interface_name
from ..utils.metering import Meter
