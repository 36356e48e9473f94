# Interpreter-start shim for the UNMODIFIED reference's AD-PSGD gossip process (a forkserver child,
# so nothing can be patched from the parent): gossip/gossiper.py:50-51 reads the legacy THD-era
# attribute `torch.distributed._backend`, which current PyTorch no longer defines.  With c10d that
# attribute was `dist_backend.UNDEFINED` (-1); restore exactly that value.
try:
    import torch.distributed as _dist
    if not hasattr(_dist, '_backend'):
        _dist._backend = -1
except Exception:       # torch missing / broken: not our business here
    pass
