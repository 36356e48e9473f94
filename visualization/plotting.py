"""
Offline analysis of training runs (parity: ``visualization/plotting.py``).

Inputs are the per-rank CSVs written by ``gossip_sgd.py`` /
``gossip_sgd_adpsgd.py`` (``{tag}out_r{rank}_n{world}.csv``: 4 header lines, then
``Epoch,itr,BT(s),...,val``) -- identical to the reference's format, so logs
from either code base can be mixed -- plus, for the Transformer figure of the
paper, fairseq training logs (their training code is not part of either repo).

    python visualization/plotting.py scaling --dir results/ --algo SGP:SGP_IB_ --nodes 4 8 16 32
    python visualization/plotting.py curves  --dir results/ --algo SGP:SGP_ --algo AR:AR_ --nodes 8 16
    python visualization/plotting.py bench   SCALE_r01.json
    python visualization/plotting.py trace   run_r0.json run_r1.json      (--trace_file output)
    python visualization/plotting.py transformer --log sgp=ps.out --log sgd=ar.out --world 8

Unlike the reference (hard-coded experiment tags and an iterations-per-epoch
table for 4/8/16/32 nodes only) everything is data-driven: runs are described
on the command line and the end-of-epoch row is the last training row of each
epoch, whatever the world size.
"""

from __future__ import annotations

import argparse
import json
import os
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import pandas as pd


def _plt(headless=True):
    import matplotlib
    if headless:
        matplotlib.use('Agg')
    matplotlib.rc('font', size=16)
    import matplotlib.pyplot as plt
    return plt


@dataclass
class RunSpec:
    label: str          # legend entry, e.g. 'SGP'
    tag: str            # file tag prefix
    directory: str = '.'

    def path(self, rank: int, world: int) -> str:
        return os.path.join(self.directory, '{}out_r{}_n{}.csv'.format(self.tag, rank, world))


# --------------------------------------------------------------------------- #
# parsing
# --------------------------------------------------------------------------- #
def read_rank_csv(path: str) -> pd.DataFrame:
    return pd.read_csv(path, skiprows=4).drop_duplicates()


def epoch_summary(df: pd.DataFrame) -> pd.DataFrame:
    """One row per epoch: training error / time-per-iteration from the last
    training row of the epoch, validation error from the ``itr == -1`` row."""
    train = df[df['itr'] >= 0]
    last = train.sort_values(['Epoch', 'itr']).groupby('Epoch').tail(1).set_index('Epoch')
    out = pd.DataFrame({
        'itr_in_epoch': last['itr'] + 1,
        'train_err': 100.0 - last['avg:Prec@1'],
        'time_per_itr': last['avg:BT(s)'],
        'nn_time_per_itr': last['avg:NT(s)'],
        'data_time_per_itr': last['avg:DT(s)'],
    })
    val = df[(df['itr'] == -1) & (df['val'] != -1)].set_index('Epoch')['val']
    out['val_err'] = 100.0 - val
    return out


def parse_csv(world_size: int, tag: str, fpath: str = None, directory: str = '.',
              ranks: Optional[Sequence[int]] = None) -> pd.DataFrame:
    """Rank-averaged per-epoch statistics with cumulative iterations and
    wall-clock (``time`` = iterations x mean time-per-iteration)."""
    per_rank = []
    for r in (ranks if ranks is not None else range(world_size)):
        path = fpath.format(tag=tag, r=r, n=world_size) if fpath else \
            RunSpec('', tag, directory).path(r, world_size)
        if os.path.isfile(path):
            per_rank.append(epoch_summary(read_rank_csv(path)))
    if not per_rank:
        raise FileNotFoundError('no CSVs for tag %r, world %d' % (tag, world_size))
    cat = pd.concat(per_rank, keys=range(len(per_rank)))
    mean = cat.groupby(level=1).mean()
    mean = mean.rename(columns={'train_err': 'train_mean', 'val_err': 'val_mean',
                                'time_per_itr': 'time_mean'})
    mean['itr'] = mean['itr_in_epoch'].cumsum()
    mean['time'] = mean['itr'] * mean['time_mean'].iloc[-1]
    mean['ranks_found'] = len(per_rank)
    return mean


_FAIRSEQ_KV = re.compile(r'(\w+)\s+([-+0-9.eE]+)')


def parse_transformer_out(world_size: int, tag: str, fpath: str, itr_scale: float = 1.0
                          ) -> pd.DataFrame:
    """fairseq logs prefixed with ``"{rank}: "``: collect per-epoch
    ``valid_nll_loss`` / ``valid_ppl`` / ``num_updates`` and the largest
    ``train_wall``; epoch 1 is skipped (warm-up), ranks are averaged."""
    rows: Dict[int, Dict[int, Dict[str, float]]] = {}
    with open(fpath.format(tag=tag)) as f:
        for line in f:
            m = re.match(r'\s*(\d+):', line)
            if not m or '|' not in line:
                continue
            rank = int(m.group(1))
            em = re.search(r'epoch\s+(\d+)', line)
            if not em:
                continue
            ep = int(em.group(1))
            if ep == 1:
                continue
            kv = {k: float(v) for k, v in _FAIRSEQ_KV.findall(line.split('|', 1)[1])}
            slot = rows.setdefault(rank, {}).setdefault(ep, {})
            if 'train_wall' in kv:
                slot['time'] = max(slot.get('time', 0.0), kv['train_wall'])
            if 'valid_nll_loss' in kv:
                slot['nll'] = kv['valid_nll_loss']
                slot['ppl'] = kv.get('valid_ppl', float('nan'))
                slot['itr'] = kv.get('num_updates', float('nan')) * itr_scale
    frames = [pd.DataFrame.from_dict(eps, orient='index') for eps in rows.values() if eps]
    if not frames:
        raise ValueError('no fairseq records in ' + fpath)
    return pd.concat(frames).groupby(level=0).mean().dropna(subset=['nll']).sort_index()


# --------------------------------------------------------------------------- #
# figures
# --------------------------------------------------------------------------- #
def plot_scaling(runs: List[RunSpec], nodes: Sequence[int], save_fname='scaling.pdf',
                 throughput=False, batch_per_node=256, headless=True):
    """Average time per iteration (or images/s) versus number of nodes."""
    plt = _plt(headless)
    fig, ax = plt.subplots()
    table = pd.DataFrame({'nodes': list(nodes)})
    for run in runs:
        ys = []
        for n in nodes:
            try:
                tpi = parse_csv(n, run.tag.format(n=n), directory=run.directory)['time_mean'] \
                    .dropna().iloc[-1]
                ys.append((batch_per_node * n) / tpi if throughput else tpi)
            except FileNotFoundError:
                ys.append(np.nan)
        table[run.label] = ys
        ax.plot(table['nodes'], ys, marker='o', label=run.label)
    ax.set_xticks(list(nodes))
    ax.set_xlabel('Number of nodes')
    ax.set_ylabel('Throughput (images/s)' if throughput else 'Time per iteration (s)')
    ax.grid(True, which='both', alpha=0.4)
    ax.legend()
    fig.tight_layout()
    fig.savefig(save_fname)
    return table


def plot_itrs(runs: List[RunSpec], nodes: Sequence[int], save_fname='itr.pdf', val=False,
              headless=True):
    """Train / validation error versus wall-clock time."""
    plt = _plt(headless)
    fig, ax = plt.subplots()
    styles = ['-', '--', ':', '-.']
    for j, n in enumerate(nodes):
        for run in runs:
            try:
                df = parse_csv(n, run.tag.format(n=n), directory=run.directory)
            except FileNotFoundError:
                continue
            ax.plot(df['time'], df['val_mean' if val else 'train_mean'],
                    linestyle=styles[j % len(styles)], label='%s %d nodes' % (run.label, n))
    ax.set_ylabel('Validation Error (%)' if val else 'Training Error (%)')
    ax.set_xlabel('Time (s)')
    ax.grid(True, which='both', alpha=0.4)
    ax.legend(prop={'size': 12})
    fig.tight_layout()
    fig.savefig(save_fname)


def plot_transformer(logs: Dict[str, str], world_size: int, save_fname='transformer.pdf',
                     headless=True):
    plt = _plt(headless)
    fig, ax = plt.subplots()
    for label, path in logs.items():
        df = parse_transformer_out(world_size, '', path)
        ax.plot(df['itr'], df['nll'], label=label)
    ax.set_ylabel('Validation Loss (NLL)')
    ax.set_xlabel('Opt. steps')
    ax.grid(True, which='both', alpha=0.4)
    ax.legend()
    fig.tight_layout()
    fig.savefig(save_fname)


def plot_bench(json_paths: Sequence[str], save_fname='bench_scaling.pdf', headless=True):
    """images/s versus GPUs from bench.py JSON lines (one object per line or a
    list); both arms ('ours' / 'reference') if present."""
    plt = _plt(headless)
    recs = []
    for p in json_paths:
        with open(p) as f:
            txt = f.read().strip()
        try:
            data = json.loads(txt)
            recs += data if isinstance(data, list) else [data]
        except json.JSONDecodeError:
            recs += [json.loads(l) for l in txt.splitlines() if l.startswith('{')]
    fig, ax = plt.subplots()
    df = pd.DataFrame([{'impl': r.get('impl', 'ours'), 'n': r['n_gpus'], 'value': r['value']}
                       for r in recs if 'value' in r])
    for impl, g in df.groupby('impl'):
        g = g.sort_values('n')
        ax.plot(g['n'], g['value'], marker='o', label=impl)
    ax.set_xlabel('GPUs')
    ax.set_ylabel('images/s')
    ax.grid(True, alpha=0.4)
    ax.legend()
    fig.tight_layout()
    fig.savefig(save_fname)
    return df


def summarize_traces(json_paths: Sequence[str]):
    """Per-rank span totals of the Chrome traces written by ``--trace_file`` (see
    ``stochastic_gradient_push_b200/utils/tracing.py``): one row per (rank, span) with call
    count, total / mean milliseconds and the share of the traced wall-clock, plus the summed
    ``exposed_comm_ms`` counter (communication the step actually waited for)."""
    rows = []
    for p in json_paths:
        with open(p) as f:
            ev = json.load(f)['traceEvents']
        spans = [e for e in ev if e.get('ph') == 'X']
        if not spans:
            continue
        rank = spans[0]['pid']
        wall = (max(e['ts'] + e['dur'] for e in spans) - min(e['ts'] for e in spans)) / 1e3
        for name in sorted({e['name'] for e in spans}):
            durs = [e['dur'] / 1e3 for e in spans if e['name'] == name]
            rows.append({'rank': rank, 'span': name, 'calls': len(durs), 'total_ms': sum(durs),
                         'mean_ms': sum(durs) / len(durs), 'share': sum(durs) / wall if wall else 0.0})
        exposed = [e['args']['exposed_comm_ms'] for e in ev
                   if e.get('ph') == 'C' and e.get('name') == 'exposed_comm_ms']
        if exposed:
            rows.append({'rank': rank, 'span': '(exposed_comm counter)', 'calls': len(exposed),
                         'total_ms': sum(exposed), 'mean_ms': sum(exposed) / len(exposed),
                         'share': sum(exposed) / wall if wall else 0.0})
    return pd.DataFrame(rows, columns=['rank', 'span', 'calls', 'total_ms', 'mean_ms', 'share'])


# --------------------------------------------------------------------------- #
def _runs(specs, directory):
    out = []
    for s in specs or []:
        label, tag = s.split(':', 1)
        out.append(RunSpec(label, tag, directory))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    sub = ap.add_subparsers(dest='cmd', required=True)
    for name in ('scaling', 'curves'):
        p = sub.add_parser(name)
        p.add_argument('--dir', default='.')
        p.add_argument('--algo', action='append', help='LABEL:TAG (TAG may contain {n})')
        p.add_argument('--nodes', nargs='+', type=int, default=[4, 8, 16, 32])
        p.add_argument('--out', default=None)
        p.add_argument('--val', action='store_true')
        p.add_argument('--throughput', action='store_true')
    p = sub.add_parser('bench')
    p.add_argument('json', nargs='+')
    p.add_argument('--out', default='bench_scaling.pdf')
    p = sub.add_parser('trace')
    p.add_argument('json', nargs='+', help='PREFIX_r<rank>.json files written by --trace_file')
    p = sub.add_parser('transformer')
    p.add_argument('--log', action='append', help='LABEL=PATH')
    p.add_argument('--world', type=int, default=8)
    p.add_argument('--out', default='transformer.pdf')
    a = ap.parse_args(argv)
    if a.cmd == 'scaling':
        print(plot_scaling(_runs(a.algo, a.dir), a.nodes, a.out or 'scaling.pdf', a.throughput))
    elif a.cmd == 'curves':
        plot_itrs(_runs(a.algo, a.dir), a.nodes, a.out or 'itr.pdf', a.val)
    elif a.cmd == 'bench':
        print(plot_bench(a.json, a.out))
    elif a.cmd == 'trace':
        print(summarize_traces(a.json).to_string(index=False, float_format=lambda v: '%.3f' % v))
    else:
        plot_transformer(dict(s.split('=', 1) for s in a.log), a.world, a.out)


if __name__ == '__main__':
    main()
