"""Offline analysis: CSV / fairseq-log parsing (no matplotlib needed)."""
import os

import pytest

from stochastic_gradient_push_b200.cli.common import CSVLog
from stochastic_gradient_push_b200.utils import Meter
from visualization import plotting


def _write_run(d, tag, world, epochs=3, itrs=5):
    for r in range(world):
        log = CSVLog(os.path.join(d, '%sout_r%d_n%d.csv' % (tag, r, world)), world, 4, 32)
        bt = Meter()
        for ep in range(epochs):
            top1 = Meter()
            loss = Meter()
            for i in range(itrs):
                bt.update(0.1 + 0.01 * r)
                top1.update(10.0 * (ep + 1) + r)
                loss.update(2.0 - 0.1 * ep)
                log.train_row(ep, i, bt, bt, bt, loss, top1, top1)
            log.val_row(ep, bt, bt, bt, 20.0 * (ep + 1))


def test_parse_csv_rank_average(tmp_path):
    _write_run(str(tmp_path), 'SGP_', 2)
    df = plotting.parse_csv(2, 'SGP_', directory=str(tmp_path))
    assert list(df.index) == [0, 1, 2] and df['ranks_found'].iloc[0] == 2
    assert abs(df['train_mean'].iloc[0] - (100 - 10.5)) < 1e-6       # mean over ranks 0,1
    assert abs(df['val_mean'].iloc[2] - 40.0) < 1e-6
    assert abs(df['time_mean'].iloc[-1] - 0.105) < 1e-3
    assert list(df['itr']) == [5, 10, 15]
    assert abs(df['time'].iloc[-1] - 15 * df['time_mean'].iloc[-1]) < 1e-9
    # reference-style explicit path template still works
    df2 = plotting.parse_csv(2, 'SGP_', fpath=str(tmp_path) + '/{tag}out_r{r}_n{n}.csv')
    assert df2.equals(df)
    with pytest.raises(FileNotFoundError):
        plotting.parse_csv(4, 'nope_', directory=str(tmp_path))


def test_parse_fairseq_log(tmp_path):
    p = tmp_path / 'transformer_ps_test.out'
    lines = []
    for rank in range(2):
        for ep in (1, 2, 3):
            lines.append('%d: | epoch %03d | loss 4.1 | train_wall %d' % (rank, ep, 100 * ep + rank))
            lines.append('%d: | epoch %03d | valid on subset | valid_nll_loss %.2f | valid_ppl %.2f '
                         '| num_updates %d' % (rank, ep, 3.0 - 0.2 * ep, 8.0 - ep, 1000 * ep))
    p.write_text('\n'.join(lines))
    df = plotting.parse_transformer_out(2, 'ps', str(tmp_path / 'transformer_{tag}_test.out'))
    assert list(df.index) == [2, 3]                   # epoch 1 skipped
    assert abs(df.loc[2, 'nll'] - 2.6) < 1e-9 and abs(df.loc[3, 'itr'] - 3000) < 1e-9
    assert abs(df.loc[2, 'time'] - 200.5) < 1e-9      # rank average of the max train_wall


def test_trace_summary(tmp_path):
    """`plotting.py trace`: per-span totals and the exposed-communication counter."""
    from stochastic_gradient_push_b200.utils import tracing
    import visualization.plotting as P
    paths = []
    for rank in range(2):
        t = tracing.Tracer(str(tmp_path / 'tr'), rank=rank, enabled=True, nvtx=False)
        for _ in range(3):
            with t.span('forward'):
                pass
            with t.span('backward'):
                pass
            t.counter('exposed_comm_ms', 1.5)
        paths.append(t.dump())
    df = P.summarize_traces(paths)
    assert set(df['rank']) == {0, 1}
    fwd = df[(df['rank'] == 1) & (df['span'] == 'forward')].iloc[0]
    assert fwd['calls'] == 3 and fwd['total_ms'] >= 0
    exp = df[(df['rank'] == 0) & (df['span'] == '(exposed_comm counter)')].iloc[0]
    assert exp['calls'] == 3 and abs(exp['total_ms'] - 4.5) < 1e-9
    P.main(['trace'] + paths)          # CLI entry prints the table
