"""Synthetic data config: a drop-in for the reference's datasets/*_config.py files (contract: datasets/multid_config.py:25-144).

    python train.py --data_config genesis_b200/datasets/synthetic_config.py --model_config <plug-in>.py ...

Importing this file registers the data flags the model configs and train.py read (`img_size`, `K_steps`, `num_workers`;
datasets/multid_config.py:32-39) and `load(cfg)` returns `(train_loader, val_loader, test_loader)` whose batches are dicts
`{'input': float32 [B,3,H,W] in [0,1], 'instances': int64 [B,1,H,W]}` (multid_config.py:131-144) -- dataset-SHAPED images from
genesis_b200/datasets/synth.py, because the real datasets cannot be downloaded here.  The loaders expose `batch_size` and
`__len__` (train.py:485-496 reads both) and are re-iterable (train.py:216 restarts the training loader every epoch).
"""
import os
import sys

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import genesis_b200  # noqa: E402

try:
    from forge import flags
except ImportError:  # reference Forge not importable (needs TensorFlow / imp): use the bundled stand-in
    genesis_b200.enable_compat()
    from forge import flags

from genesis_b200.datasets import synth  # noqa: E402

flags.DEFINE_string('synthetic_kind', 'multid', '{multid, stacks, rooms} - which dataset the images are shaped like.')
flags.DEFINE_integer('synthetic_size', 256, 'Number of images per split.')
flags.DEFINE_boolean('synthetic_test_split', False, 'Also return a test loader (train.py then runs the FID stage, which '
                                                    'needs downloaded Inception weights).')
flags.DEFINE_boolean('load_instances', True, 'Load instances.')
flags.DEFINE_integer('img_size', 64, 'Dimension of images. Images are square.')
flags.DEFINE_integer('num_workers', 0, 'Number of threads for loading data (unused: images are generated in memory).')
flags.DEFINE_integer('K_steps', 5, 'Number of recurrent steps.')


class SyntheticLoader(object):
    """Minimal torch DataLoader look-alike over an in-memory split: `batch_size`, `__len__`, re-iterable, optional
    per-epoch shuffle with its OWN numpy generator (train.py:128 makes CUDA the default tensor type, under which
    torch.utils.data's RandomSampler cannot build its CPU permutation)."""

    def __init__(self, images, instances, batch_size, shuffle, seed):
        self.images, self.instances = images, instances
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.rng = np.random.RandomState(seed)
        self.dataset = self              # DataLoader.dataset, for callers that ask len(loader.dataset)

    def __len__(self):
        return (self.images.shape[0] + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.images.shape[0]
        order = self.rng.permutation(n) if self.shuffle else np.arange(n)
        for i in range(0, n, self.batch_size):
            idx = order[i:i + self.batch_size]
            batch = {'input': torch.from_numpy(self.images[idx]).to('cpu')}
            if self.instances is not None:
                batch['instances'] = torch.from_numpy(self.instances[idx]).to('cpu')
            yield batch


def load(cfg, **unused_kwargs):
    """-> (train_loader, val_loader, test_loader); test_loader is None unless --synthetic_test_split."""
    del unused_kwargs
    gen = synth.GENERATORS[cfg.synthetic_kind]
    n = int(cfg.synthetic_size)
    want_inst = getattr(cfg, 'load_instances', True)
    loaders = []
    for split, seed in (('train', 1), ('val', 2), ('test', 3)):
        if split == 'test' and not getattr(cfg, 'synthetic_test_split', False):
            loaders.append(None)
            continue
        x, inst = gen(n, cfg.img_size, seed)
        loaders.append(SyntheticLoader(x, inst if want_inst else None, cfg.batch_size, shuffle=(split == 'train'), seed=seed))
    return tuple(loaders)
