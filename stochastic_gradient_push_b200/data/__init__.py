"""Input pipeline: pre-decoded uint8 shards + GPU-side augmentation (``data/shards.py``)."""
from .shards import (GpuAugment, ShardLoader, write_shards, write_shards_from_arrays,  # noqa: F401
                     IMAGENET_MEAN, IMAGENET_STD)
