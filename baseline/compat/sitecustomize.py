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

# Do not shadow another sitecustomize further down sys.path (e.g. a measurement harness's
# import hook): run the next one found, exactly as the interpreter would have without this file.
try:
    import os as _os
    import sys as _sys
    _here = _os.path.dirname(_os.path.abspath(__file__))
    for _p in list(_sys.path):
        _cand = _os.path.join(_p or '.', 'sitecustomize.py')
        if _os.path.abspath(_p or '.') != _here and _os.path.isfile(_cand):
            with open(_cand) as _f:
                exec(compile(_f.read(), _cand, 'exec'), {'__name__': 'sitecustomize', '__file__': _cand})
            break
except Exception:
    pass
