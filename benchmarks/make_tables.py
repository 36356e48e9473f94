#!/usr/bin/env python
"""Print markdown rows from bench.py JSON lines:  python benchmarks/make_tables.py gpurun_out/r2c6_*.json"""
import json
import sys


def last_json(path):
    try:
        ls = [l for l in open(path) if l.startswith('{')]
        return json.loads(ls[-1]) if ls else None
    except Exception:
        return None


def main():
    print('| file | impl | algo | N | batch/GPU | img/s (device) | ms/step | img/s (e2e) | exposed comm ms | note |')
    print('|---|---|---|---|---|---|---|---|---|---|')
    for p in sys.argv[1:]:
        d = last_json(p)
        name = p.split('/')[-1]
        if d is None:
            print('| %s | – | – | – | – | – | – | – | – | no JSON line |' % name)
            continue
        if 'unavailable' in d:
            print('| %s | reference | – | – | – | – | – | – | – | unavailable: %s |' % (name, d['unavailable'][:80]))
            continue
        cfg = d.get('config', {})
        ex = (d.get('exposed_comm') or {}).get('ms_per_step')
        e2e = (d.get('e2e') or {}).get('value')
        note = cfg.get('graph', cfg.get('algorithm', ''))
        if d.get('gossip_rounds_completed_rank0') is not None:
            note += '; %d bilateral rounds (rank 0)' % d['gossip_rounds_completed_rank0']
        print('| %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |' % (
            name, d.get('impl', '?'), d.get('metric', '').replace('resnet50_', '').replace('_images_per_sec', ''),
            d.get('n_gpus'), cfg.get('per_gpu_batch'), d.get('value'), d.get('ms_per_step'), e2e,
            '–' if ex is None else ex, note))


if __name__ == '__main__':
    main()
