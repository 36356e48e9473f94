"""Decode an ImageFolder tree into uint8 shards once:

    python -m stochastic_gradient_push_b200.data.make_shards /data/imagenet/train /data/imagenet_shards/train
    python -m stochastic_gradient_push_b200.data.make_shards /data/imagenet/val   /data/imagenet_shards/val

then train with ``--dataset_dir /data/imagenet_shards --data_format shards``.
(ImageNet-1k at 256x256: 1.28 M x 196 KB = 252 GB; the maps are paged in on demand.)
"""
import argparse
import json

from .shards import write_shards


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('image_folder')
    ap.add_argument('out_dir')
    ap.add_argument('--size', type=int, default=256, help='stored square size (short side resize + centre crop)')
    ap.add_argument('--shard_size', type=int, default=4096, help='images per shard file')
    ap.add_argument('--limit', type=int, default=None, help='only the first N files (smoke tests)')
    args = ap.parse_args(argv)
    index = write_shards(args.image_folder, args.out_dir, args.size, args.shard_size, args.limit)
    print(json.dumps({'shards': len(index['counts']), 'images': sum(index['counts']), 'size': index['size'],
                      'classes': len(index['classes'])}))


if __name__ == '__main__':
    main()
