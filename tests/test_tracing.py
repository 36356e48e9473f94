"""utils/tracing.py: Chrome-trace recorder (and NVTX mirror on CUDA machines)."""
import json
import threading
import time

from stochastic_gradient_push_b200.utils import tracing


def test_disabled_tracer_is_a_noop():
    tracing.disable()
    t = tracing.get_tracer()
    assert not t.enabled
    s1, s2 = tracing.span('a'), tracing.span('b', x=1)
    assert s1 is s2                      # shared null context: nothing allocated per call
    with s1:
        pass
    tracing.counter('c', 1.0)
    tracing.instant('i')
    assert t.events == [] and t.dump() is None


def test_spans_counters_and_dump(tmp_path):
    t = tracing.enable(str(tmp_path / 'tr'), rank=3, nvtx=False)
    try:
        with tracing.span('outer', itr=7):
            with tracing.span('inner'):
                time.sleep(0.01)
        tracing.counter('exposed_comm_ms', 2.5)
        tracing.instant('checkpoint', epoch=1)

        def worker():
            with tracing.span('thread-span'):
                pass
        th = threading.Thread(target=worker)
        th.start()
        th.join()
    finally:
        assert tracing.disable() is t
    summ = t.summary()
    assert summ['outer'][0] == 1 and summ['inner'][1] >= 9.0 and summ['outer'][1] >= summ['inner'][1]
    path = t.dump()
    assert path.endswith('tr_r3.json')
    doc = json.load(open(path))
    ev = doc['traceEvents']
    assert ev[0]['ph'] == 'M' and ev[0]['args']['name'] == 'rank 3'
    by = {e['name']: e for e in ev if e['ph'] != 'M'}
    assert by['outer']['args'] == {'itr': 7} and by['outer']['pid'] == 3
    inner, outer = by['inner'], by['outer']
    assert outer['ts'] <= inner['ts'] and inner['ts'] + inner['dur'] <= outer['ts'] + outer['dur'] + 1
    assert by['exposed_comm_ms']['ph'] == 'C' and by['exposed_comm_ms']['args'] == {'exposed_comm_ms': 2.5}
    assert by['thread-span']['tid'] != outer['tid']


def test_event_cap():
    t = tracing.Tracer(enabled=True, nvtx=False, max_events=3)
    for i in range(10):
        with t.span('s%d' % i):
            pass
    assert len(t.events) == 3
