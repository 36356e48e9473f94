"""
Running statistics meter (API parity: ``gossip/utils/metering.py:13-80``,
duplicated in the reference as ``experiment_utils/metering.py``).

val / avg / sum / count / std (sample std from running sums) and, when
``stateful``, the full history plus the mean absolute deviation.  All
statistics are plain attributes refreshed on every ``update`` because
checkpoints store ``meter.__dict__`` verbatim and re-hydrate a meter from it
(``gossip_sgd.py:214-216, 257-259``).
"""


class Meter(object):

    def __init__(self, init_dict=None, ptag='Time', stateful=False,
                 csv_format=True):
        self.reset()
        self.ptag = ptag
        self.stateful = stateful
        self.value_history = [] if stateful else None
        self.csv_format = csv_format
        if init_dict is not None:
            for key, value in init_dict.items():
                setattr(self, key, value)

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0
        self.std = 0
        self.sqsum = 0
        self.mad = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
        self.sqsum += (val ** 2) * n
        if self.count > 1:
            var = (self.sqsum - (self.sum ** 2) / self.count) / (self.count - 1)
            self.std = max(var, 0.0) ** 0.5
        if self.stateful:
            self.value_history.append(val)
            avg = self.avg
            self.mad = sum(abs(v - avg) for v in self.value_history) \
                / len(self.value_history)

    def __str__(self):
        spread = self.mad if self.stateful else self.std
        if self.csv_format:
            return '{:.3f},{:.3f},{:.3f}'.format(self.val, self.avg, spread)
        return '{}: {:.3f} ({:.3f} +- {:.3f})'.format(
            self.ptag, self.val, self.avg, spread)
