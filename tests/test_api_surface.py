"""SURVEY.md 2.6 as an executable checklist: every public name, constructor argument (in the
reference's positional order), method and attribute a user of the reference relies on exists here
under the same name.  Pure introspection plus one single-rank CPU instance -- no GPU, no peers."""
import inspect

import pytest
import torch


def _params(fn):
    return [p for p in inspect.signature(fn).parameters if p != 'self']


def _has_prefix(fn, names):
    got = _params(fn)
    assert got[:len(names)] == list(names), (fn, got)


def test_drop_in_package_exports():
    import gossip
    for name in ('BilatGossipDataParallel', 'GossipDataParallel', 'GraphManager', 'MixingManager',
                 'UniformMixing', 'PushSum', 'PushPull', 'DynamicDirectedExponentialGraph',
                 'NPeerDynamicDirectedExponentialGraph', 'DynamicBipartiteExponentialGraph',
                 'DynamicDirectedLinearGraph', 'DynamicBipartiteLinearGraph', 'RingGraph'):
        assert hasattr(gossip, name), name
    from gossip.gossiper import BilatPushPull, Gossiper          # not re-exported by the reference either
    assert issubclass(BilatPushPull, Gossiper)
    from gossip.utils import (flatten_tensors, unflatten_tensors, group_by_dtype, communicate,   # noqa: F401
                              make_logger, is_power_of, create_process_group)
    from gossip.utils.metering import Meter
    _has_prefix(Meter.__init__, ['init_dict', 'ptag', 'stateful', 'csv_format'])
    import experiment_utils
    for name in ('ClusterManager', 'Meter', 'make_logger', 'get_tcp_interface_name'):
        assert hasattr(experiment_utils, name), name


def test_gossip_data_parallel_signature_and_methods():
    from gossip import GossipDataParallel as GDP
    _has_prefix(GDP.__init__, ['module', 'device_ids', 'rank', 'world_size', 'graph', 'mixing', 'comm_device',
                               'push_sum', 'overlap', 'synch_freq', 'verbose', 'use_streams',
                               'nprocs_per_node', 'local_node_group'])
    d = {k: v.default for k, v in inspect.signature(GDP.__init__).parameters.items()}
    assert d['push_sum'] is True and d['overlap'] is False and d['synch_freq'] == 0
    assert d['use_streams'] is True and d['nprocs_per_node'] == 1 and d['verbose'] is False
    for m in ('forward', 'transfer_params', 'sync_comms', 'block', 'state_dict', 'load_state_dict', 'train',
              'eval', 'update_gossiper', 'ps_numerator', 'unbias', 'scatter', 'parallel_apply', 'gather',
              '_query_gossip_queue'):
        assert callable(getattr(GDP, m)), m
    _has_prefix(GDP.transfer_params, ['mix'])
    _has_prefix(GDP.state_dict, ['finish_gossip'])
    _has_prefix(GDP.load_state_dict, ['load_dict'])
    _has_prefix(GDP.update_gossiper, ['attr', 'val'])
    _has_prefix(GDP._query_gossip_queue, ['non_blocking'])


def test_bilat_gossip_data_parallel_signature_and_methods():
    from gossip import BilatGossipDataParallel as B
    _has_prefix(B.__init__, ['module', 'device_ids', 'master_addr', 'master_port', 'backend', 'world_size',
                             'rank', 'graph_class', 'mixing_class', 'num_peers', 'comm_device', 'lr',
                             'momentum', 'weight_decay', 'nesterov', 'verbose', 'network_interface_type',
                             'tcp_interface_name'])
    d = {k: v.default for k, v in inspect.signature(B.__init__).parameters.items()}
    assert (d['lr'], d['momentum'], d['weight_decay'], d['nesterov'], d['num_peers']) == (0.1, 0.9, 1e-4, True, 1)
    for m in ('forward', 'update_lr', 'enable_gossip', 'disable_gossip', 'block', 'sync_comms', 'train', 'eval',
              '_pull_model', '_transfer_grads', 'communicator_warmup'):
        assert callable(getattr(B, m)), m


def test_gossiper_family_signature():
    from gossip.gossiper import Gossiper, PushSum, PushPull, BilatPushPull
    for cls in (Gossiper, PushSum, PushPull, BilatPushPull):
        _has_prefix(cls.__init__, ['msg', 'graph', 'device', 'mixing', 'logger', 'rank', 'world_size'])
        for m in ('refresh_peers_', 'refresh_mixing_weights_', 'mix_out_msg_', 'clean_msg_buffers_',
                  'parse_in_msg_buffer', 'mix'):
            assert callable(getattr(cls, m)), (cls, m)
        assert isinstance(inspect.getattr_static(cls, 'ps_weight'), property)
        assert isinstance(inspect.getattr_static(cls, 'peers_per_itr'), property)
    _has_prefix(PushSum.mix, ['out_msg', 'ps_weight', 'residual'])
    _has_prefix(PushPull.mix, ['out_msg', 'ps_weight', 'residual'])
    _has_prefix(BilatPushPull.mix, ['out_msg'])
    _has_prefix(Gossiper.refresh_peers_, ['rotate'])
    _has_prefix(Gossiper.refresh_mixing_weights_, ['residual_adjusted'])
    _has_prefix(Gossiper.parse_in_msg_buffer, ['residual'])


def test_graph_and_mixing_manager_surface():
    import gossip
    from stochastic_gradient_push_b200.topology import Edge
    _has_prefix(gossip.GraphManager.__init__, ['rank', 'world_size', 'nprocs_per_node', 'local_rank', 'peers_per_itr'])
    g = gossip.NPeerDynamicDirectedExponentialGraph(0, 8, peers_per_itr=1)
    for m in ('get_peers', 'get_edges', 'is_regular_graph', 'is_bipartite_graph', 'is_passive', 'is_dynamic_graph'):
        assert callable(getattr(g, m)), m
    _has_prefix(g.get_peers, ['rotate'])
    _has_prefix(g.get_edges, ['rotate'])
    out, ins = g.get_peers()
    assert out == [1] and ins == [7]
    g.peers_per_itr = 2
    assert g.peers_per_itr == 2 and len(g.get_peers()[0]) == 2
    e = g.phone_book[0][0]
    assert isinstance(e, Edge) and {'src', 'dest', 'process_group'} <= set(dir(e))
    mm = gossip.UniformMixing(g, torch.device('cpu'))
    assert isinstance(mm, gossip.MixingManager) and mm.is_regular() and mm.is_uniform()
    w = mm.get_mixing_weights(residual_adjusted=True)
    assert 'lo' in w and 'uniform' in w and float(w['uniform']) == 1.0
    w = mm.get_mixing_weights(residual_adjusted=False)
    assert abs(float(w['lo']) - 1.0 / 3) < 1e-6 and abs(float(w['uniform']) - 1.0 / 3) < 1e-6


def test_gossip_data_parallel_instance_attributes():
    """the attributes the reference's CLI and users touch, on a single-rank CPU instance"""
    import gossip
    net = torch.nn.Linear(4, 2)
    m = gossip.GossipDataParallel(net, graph=gossip.NPeerDynamicDirectedExponentialGraph(0, 1),
                                  rank=0, world_size=1)
    for a in ('module', 'ps_weight', 'is_ps_numerator', 'gossip_enable', 'gossiping', 'params_mixed', 'overlap',
              'synch_freq', 'asynch', 'lazy_mixing', 'num_updates', 'dist_config', 'gossip_stream',
              'device_ids', 'output_device', 'logger'):
        assert hasattr(m, a), a
    for k in ('verbose', 'comm_device', 'graph', 'mixing', 'push_sum', 'rank', 'process_rank', 'world_size',
              'cpu_comm', 'gossipers'):
        assert k in m.dist_config, k
    assert m.module is net
    sd = m.state_dict()
    assert set(sd) == {'state_dict', 'ps_weight', 'is_ps_numerator'}
    m.load_state_dict(sd)
    m.train()
    y = m(torch.randn(3, 4))
    assert y.shape == (3, 2)
    m.eval()
    m.update_gossiper('peers_per_itr', 1)
    m.block()
    m.sync_comms()


@pytest.mark.parametrize('script', ['gossip_sgd', 'gossip_sgd_adpsgd'])
def test_cli_flags_and_registries(script):
    """every flag of the reference scripts is accepted, and the integer registries exist"""
    import importlib
    mod = importlib.import_module('stochastic_gradient_push_b200.cli.' + script)
    common = importlib.import_module('stochastic_gradient_push_b200.cli.common')
    reg = mod if hasattr(mod, 'GRAPH_TOPOLOGIES') else common
    assert set(reg.GRAPH_TOPOLOGIES) >= {0, 1, 2, 3, 4, 5} and set(reg.MIXING_STRATEGIES) >= {0}
    src = inspect.getsource(mod) + inspect.getsource(common)
    common_flags = ['--all_reduce', '--backend', '--batch_size', '--checkpoint_all', '--checkpoint_dir',
                    '--dataset_dir', '--graph_type', '--lr', '--master_port', '--mixing_strategy', '--momentum',
                    '--nesterov', '--network_interface_type', '--num_dataloader_workers', '--num_epochs',
                    '--overlap', '--peers_per_itr_schedule', '--print_freq', '--push_sum', '--resume',
                    '--schedule', '--seed', '--synch_freq', '--tag', '--train_fast', '--verbose', '--warmup',
                    '--weight_decay']
    if script == 'gossip_sgd':          # gossip_sgd.py:35-160
        flags = common_flags + ['--no_cuda_streams', '--num_iterations_per_training_epoch', '--num_itr_ignore',
                                '--overwrite_checkpoints']
    else:                               # gossip_sgd_adpsgd.py:36-140
        flags = common_flags + ['--bilat', '--bs_fpath', '--shared_fpath']
    for f in flags:
        assert ("'%s'" % f) in src or ('"%s"' % f) in src, f
    for fn in ('main', 'train', 'validate', 'parse_args'):
        assert callable(getattr(mod, fn, None)), fn


def test_reference_module_paths_import():
    """`from gossip.graph_manager import ...`-style imports of reference users keep working"""
    import importlib
    for name in ('gossip.distributed', 'gossip.ad_psgd', 'gossip.gossiper', 'gossip.graph_manager',
                 'gossip.mixing_manager', 'gossip.utils.helpers', 'gossip.utils.metering',
                 'experiment_utils.cluster_manager', 'experiment_utils.helpers', 'experiment_utils.metering'):
        importlib.import_module(name)
    from gossip.graph_manager import Edge, GraphManager            # noqa: F401
    from gossip.distributed import GossipDataParallel              # noqa: F401
    from gossip.ad_psgd import BilatGossipDataParallel             # noqa: F401
    from gossip.mixing_manager import MixingManager, UniformMixing  # noqa: F401
