"""data/shards.py: shard writer, rank / epoch sharding of ShardLoader, GPU-side augmentation (run on
the CPU here -- the code is device-agnostic), and the CLI with --data_format shards."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from stochastic_gradient_push_b200.data import (GpuAugment, ShardLoader, IMAGENET_MEAN, IMAGENET_STD,
                                                write_shards, write_shards_from_arrays)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_shards(path, n=50, size=32, classes=7, shard_size=16, seed=0):
    rng = np.random.RandomState(seed)
    imgs = rng.randint(0, 256, size=(n, size, size, 3), dtype=np.uint8)
    imgs[:, 0, 0, 0] = np.arange(n) % 256          # tag every image with its index
    labels = np.arange(n) % classes
    write_shards_from_arrays(str(path), imgs, labels, shard_size=shard_size)
    return imgs, labels


def test_write_shards_from_an_image_folder(tmp_path):
    from PIL import Image
    rng = np.random.RandomState(1)
    for c in ('cat', 'dog'):
        os.makedirs(str(tmp_path / 'src' / c))
        for i in range(3):
            w, h = (40, 30) if i % 2 else (24, 48)
            Image.fromarray(rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)).save(
                str(tmp_path / 'src' / c / ('%d.png' % i)))
    index = write_shards(str(tmp_path / 'src'), str(tmp_path / 'out'), store_size=16, shard_size=4)
    assert index['classes'] == ['cat', 'dog'] and index['counts'] == [4, 2] and index['size'] == 16
    loader = ShardLoader(str(tmp_path / 'out'), batch_size=6, shuffle=False, drop_last=False, pin=False)
    (x, y), = list(loader)
    assert x.shape == (6, 16, 16, 3) and x.dtype == torch.uint8 and y.tolist() == [0, 0, 0, 1, 1, 1]


def test_ranks_get_disjoint_shares_and_epochs_reshuffle(tmp_path):
    imgs, labels = _fake_shards(tmp_path)
    world, bs = 3, 4
    per_epoch = []
    for epoch in (0, 1):
        seen = []
        for r in range(world):
            loader = ShardLoader(str(tmp_path), bs, world, r, seed=5, pin=False)
            loader.set_epoch(epoch)
            assert len(loader) == (50 // world) // bs
            ids = []
            for x, y in loader:
                assert x.shape == (bs, 32, 32, 3)
                tags = x[:, 0, 0, 0].tolist()
                assert [labels[t] for t in tags] == y.tolist()       # images and labels stay paired
                ids += tags
            assert len(ids) == len(loader) * bs
            seen.append(ids)
        flat = [i for ids in seen for i in ids]
        assert len(set(flat)) == len(flat)                             # disjoint across ranks
        per_epoch.append(seen)
    assert per_epoch[0] != per_epoch[1]                                # reshuffled
    again = ShardLoader(str(tmp_path), bs, world, 0, seed=5, pin=False)
    again.set_epoch(0)
    assert [x[:, 0, 0, 0].tolist() for x, _ in again] == \
        [per_epoch[0][0][i:i + bs] for i in range(0, len(per_epoch[0][0]), bs)]     # deterministic


def test_validation_loader_covers_everything_in_order(tmp_path):
    imgs, labels = _fake_shards(tmp_path, n=37)
    loader = ShardLoader(str(tmp_path), 8, 1, 0, shuffle=False, drop_last=False, pin=False)
    xs, ys = zip(*list(loader))
    assert [len(x) for x in xs] == [8, 8, 8, 8, 5]
    assert torch.cat(ys).tolist() == labels.tolist()
    np.testing.assert_array_equal(torch.cat(xs).numpy(), imgs)


def test_abandoned_iterator_does_not_leak_a_blocked_thread(tmp_path):
    import threading
    _fake_shards(tmp_path, n=200)
    before = threading.active_count()
    it = iter(ShardLoader(str(tmp_path), 4, 1, 0, pin=False, prefetch=2))
    next(it)
    it.close()
    import time
    time.sleep(0.3)
    assert threading.active_count() <= before + 1


def test_gpu_augment_eval_is_a_normalised_centre_crop():
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 256, (3, 32, 32, 3), generator=g, dtype=torch.uint8)
    aug = GpuAugment(out_size=16, eval_ratio=0.5)
    y = aug(x, train=False)
    assert y.shape == (3, 3, 16, 16) and y.is_contiguous(memory_format=torch.channels_last)
    crop = x[:, 8:24, 8:24, :].permute(0, 3, 1, 2).float() / 255.0
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    torch.testing.assert_close(y, (crop - mean) / std, rtol=1e-4, atol=1e-4)


def test_gpu_augment_train_boxes_flips_and_determinism():
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, 256, (64, 16, 16, 3), generator=g, dtype=torch.uint8)
    full = GpuAugment(out_size=16, scale=(1.0, 1.0), ratio=(1.0, 1.0), seed=3)      # box = whole image
    y = full(x, train=True)
    plain = GpuAugment(out_size=16, eval_ratio=1.0)(x, train=False)                   # identity sampling
    flipped = plain.flip(-1)
    same = [(torch.allclose(y[i], plain[i], atol=1e-4), torch.allclose(y[i], flipped[i], atol=1e-4))
            for i in range(len(x))]
    assert all(a or b for a, b in same)
    n_flip = sum(1 for a, b in same if b and not a)
    assert 10 < n_flip < 54                                         # about half of them mirrored
    # crop boxes stay inside the image and honour scale / ratio
    aug = GpuAugment(out_size=8, seed=11)
    cx, cy, w, h, flip = aug.sample_boxes(2000)
    assert (cx - w / 2).min() >= -1e-6 and (cx + w / 2).max() <= 1 + 1e-6
    assert (cy - h / 2).min() >= -1e-6 and (cy + h / 2).max() <= 1 + 1e-6
    area = w * h
    assert area.min() >= 0.08 * 0.5 and area.max() <= 1.0 + 1e-6
    assert set(flip.tolist()) == {-1.0, 1.0}
    a, b = GpuAugment(out_size=8, seed=4), GpuAugment(out_size=8, seed=4)
    torch.testing.assert_close(a(x[:8]), b(x[:8]))                  # seeded
    assert GpuAugment(out_size=8, dtype=torch.bfloat16)(x[:2]).dtype == torch.bfloat16


def test_cli_trains_from_shards_on_two_cpu_ranks(tmp_path, master_port):
    for split, n in (('train', 96), ('val', 24)):
        _fake_shards(tmp_path / 'data' / split, n=n, size=20, classes=10, shard_size=32, seed=len(split))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(master_port),
           os.path.join(ROOT, 'gossip_sgd.py'),
           '--device', 'cpu', '--backend', 'gloo', '--model', 'tiny', '--num_classes', '10', '--image_size', '16',
           '--dataset_dir', str(tmp_path / 'data'), '--data_format', 'shards', '--batch_size', '8',
           '--verbose', 'False', '--print_freq', '2', '--amp', 'False', '--num_dataloader_workers', '0',
           '--lr', '0.05', '--push_sum', 'True', '--graph_type', '5', '--num_epochs', '2',
           '--checkpoint_dir', str(tmp_path / 'ckpt') + '/', '--num_itr_ignore', '0']
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                         env=dict(os.environ, OMP_NUM_THREADS='1'))
    assert out.returncode == 0, out.stdout[-3000:]
    for r in range(2):
        rows = open(str(tmp_path / 'ckpt' / ('out_r%d_n2.csv' % r))).read().splitlines()[5:]
        train_rows = [x.split(',') for x in rows if x.split(',')[1] != '-1']
        assert len(train_rows) >= 6 and all(float(x[12]) == float(x[12]) for x in train_rows)   # 6 iterations / epoch, finite loss
        assert any(x.split(',')[1] == '-1' for x in rows)                                        # validation ran


def test_native_row_gather_matches_numpy(tmp_path):
    try:
        from stochastic_gradient_push_b200.ops import native
        C = native.load()
    except Exception as e:
        pytest.skip('native extension unavailable: %s' % e)
    imgs, _ = _fake_shards(tmp_path, n=300, size=24, shard_size=64)
    a = ShardLoader(str(tmp_path), 32, 1, 0, seed=2, pin=False)
    b = ShardLoader(str(tmp_path), 32, 1, 0, seed=2, pin=False)
    b.use_native = False
    assert a._native() is not None
    for (xa, ya), (xb, yb) in zip(a, b):
        assert torch.equal(xa, xb) and torch.equal(ya, yb)
    # bounds are checked before any byte is copied
    out = torch.empty(2, 24 * 24 * 3, dtype=torch.uint8)
    with pytest.raises(RuntimeError):
        C.gather_rows_u8(a._ptrs, a.counts, a._row_bytes, torch.tensor([0, 99]), torch.tensor([0, 0]), out, 2)
    with pytest.raises(RuntimeError):
        C.gather_rows_u8(a._ptrs, a.counts, a._row_bytes, torch.tensor([0, 0]), torch.tensor([0, 64]), out, 2)
