"""Experiment helpers (parity: ``experiment_utils/helpers.py:18-67``)."""

from __future__ import annotations

import os
import subprocess

from ..utils.helpers import make_logger  # noqa: F401  (re-exported)

_PREFIXES = {
    'ethernet': ('ens', 'eth', 'enp', 'eno'),
    'infiniband': ('ib',),
}


def _interfaces_up():
    """Names of the links that are administratively up (``ip link show up``);
    falls back to ``/sys/class/net/*/operstate`` when ``ip`` is missing."""
    try:
        out = subprocess.run(['ip', 'link', 'show', 'up'], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, timeout=5).stdout.decode('utf-8')
        return out
    except Exception:
        names = []
        for n in os.listdir('/sys/class/net'):
            try:
                with open('/sys/class/net/%s/operstate' % n) as f:
                    if f.read().strip() in ('up', 'unknown'):
                        names.append(n)
            except OSError:
                pass
        return ' '.join(names)


def get_tcp_interface_name(network_interface_type='ethernet'):
    """First interface of the requested kind that is up (used to pin
    ``NCCL_SOCKET_IFNAME`` / ``GLOO_SOCKET_IFNAME`` on multi-NIC hosts)."""
    prefixes = _PREFIXES[network_interface_type]
    candidates = sorted(os.listdir('/sys/class/net'))
    up = _interfaces_up()
    for name in candidates:
        if name.startswith(prefixes) and name in up:
            print('Using network interface {}'.format(name))
            return name
    print('List of network interfaces found:', candidates)
    print('Prefix list being used to search:', prefixes)
    raise Exception('No proper {} interface found'.format(network_interface_type))
