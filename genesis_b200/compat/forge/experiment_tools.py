"""Forge experiment-tools stand-in: the surface the reference's callers use (train.py, scripts/*.py, datasets/*_config.py).

The real Forge (forge/forge/experiment_tools.py) imports TensorFlow, `imp` and simplejson at module level, none of which
exist in this image or on the GPU box, so `train.py` cannot import it.  This file re-implements, on the standard library
only, the functions those callers reach -- same names, arguments, return values and on-disk layout:

    load(path, *args)                 experiment_tools.py:246-258   import a config file BY PATH, call its load()
    init_checkpoint(dir, data, model, resume)      :129-230         numbered run folders, flags.json, config copies, resume
    parse_flags / assert_all_flags_parsed / print_flags             :295-338
    json_store / json_load, find_model_files, fprint                :48-60, 237-243, 444-457
    load_from_checkpoint(dir, itr, mode='torch')   :63-126          (torch branch only)
    EXPERIMENT_FOLDER, FPRINT_FILE, FLAG_FILE, _flags

Semantics worth keeping exact (train.py relies on them): `parse_flags()` consumes the flags it knows from sys.argv and
leaves the rest there for the next call (config files register their flags when `init_checkpoint` imports them), and a
value set programmatically before a re-parse survives it (train.py:101-106 sets `config.num_workers` before the data
config has defined the flag).
"""
from __future__ import print_function

import datetime
import importlib
import importlib.util
import json
import os
import re
import shutil
import subprocess
import sys

from . import flags as _flags  # noqa: F401  (train.py:142 reads fet._flags.FLAGS.__flags)

FLAG_FILE = 'flags.json'
GIT_DIFF_FILE = 'git_diff.txt'
FPRINT_FILE = 'fprint.txt'
EXPERIMENT_FOLDER = None


# ------------------------------------------------------------------------------------------------- json
def json_store(path, data):
    with open(path, 'w') as f:
        json.dump(data, f, indent=4, sort_keys=True, default=str)


def json_load(path):
    with open(path, 'r') as f:
        return json.load(f)


# ------------------------------------------------------------------------------------------------- config import
def _import_module(module_path_or_name):
    """Import a module from a file path (module name = file basename; an already imported module of that name is reused,
    as `imp.load_source` + the sys.modules check of the real Forge do) or by dotted name."""
    if module_path_or_name.endswith('.py'):
        if not os.path.exists(module_path_or_name):
            raise RuntimeError('File {} does not exist.'.format(module_path_or_name))
        name = os.path.basename(os.path.splitext(module_path_or_name)[0])
        if name in sys.modules:
            return sys.modules[name], name
        spec = importlib.util.spec_from_file_location(name, os.path.abspath(module_path_or_name))
        module = importlib.util.module_from_spec(spec)
        sys.modules[name] = module
        try:
            spec.loader.exec_module(module)
        except BaseException:
            sys.modules.pop(name, None)
            raise
        return module, name
    module = importlib.import_module(module_path_or_name)
    return module, module_path_or_name.split('.')[-1]


import_by_path = lambda path: _import_module(path)[0]  # noqa: E731  (kept: used by this repo's tests)


def load(conf_path, *args, **kwargs):
    """Loads a config: imports the file and calls its module-level `load(*args, **kwargs)`."""
    module, _ = _import_module(conf_path)
    try:
        load_func = module.load
    except AttributeError:
        raise ValueError("The config file should specify 'load' function but no such function was "
                         "found in {}".format(module.__file__))
    print("Loading '{}' from {}".format(module.__name__, module.__file__))
    parse_flags()
    return load_func(*args, **kwargs)


def _load_flags(*config_paths):
    """Importing a config file registers the flags it defines."""
    for config_path in config_paths:
        print('loading flags from', config_path)
        _import_module(config_path)


# ------------------------------------------------------------------------------------------------- flags
def parse_flags():
    """Parse the flags known so far from sys.argv; unknown arguments stay in sys.argv for a later call.  Values already
    present (parsed earlier or set programmatically) win over the re-parse, as in the real Forge."""
    f = _flags.FLAGS
    old_flags = f.__dict__['__flags'].copy()
    passthrough = f._parse_flags(args=sys.argv[1:])
    sys.argv[1:] = passthrough
    f.__dict__['__flags'].update(old_flags)
    return f.__dict__['__flags']


def _restore_flags(flags):
    _flags.FLAGS.__dict__['__flags'].update(flags)
    _flags.FLAGS.__dict__['__parsed'] = True


def print_flags():
    flags = _flags.FLAGS.__dict__['__flags']
    print('Flags:')
    print('=' * 60)
    for k in sorted(flags.keys()):
        print('\t{}: {}'.format(k, flags[k]))
    print('=' * 60)


def assert_all_flags_parsed():
    not_parsed = [a for a in sys.argv[1:] if a.startswith('--')]
    if not_parsed:
        raise RuntimeError('Failed to parse following flags: {}'.format(not_parsed))


def get_git_revision_hash():
    return subprocess.check_output(['git', 'rev-parse', 'HEAD'], stderr=subprocess.DEVNULL).strip().decode()


# ------------------------------------------------------------------------------------------------- run folders
def find_model_files(model_dir):
    """{iteration: path} of the files named `*.ckpt-<number>` in model_dir."""
    pattern = re.compile(r'.ckpt-[0-9]+$')
    files = [f.replace('.index', '') for f in os.listdir(model_dir)]
    files = [f for f in files if pattern.search(f)]
    return {int(f.split('-')[-1].split('.')[0]): os.path.join(model_dir, f) for f in files}


def init_checkpoint(checkpoint_dir, data_config, model_config, resume):
    """Create (or, with resume, find) the numbered run folder under checkpoint_dir, register + parse the flags of both config
    files, store flags.json and copies of the config files (or restore the stored flags and find the newest
    `model.ckpt-<iter>`).  Returns (experiment_folder, resume_checkpoint or None)."""
    global EXPERIMENT_FOLDER
    if not os.path.exists(checkpoint_dir):
        if resume:
            raise ValueError("Can't resume when the checkpoint dir '{}' doesn't exist.".format(checkpoint_dir))
        os.makedirs(checkpoint_dir)
    elif not os.path.isdir(checkpoint_dir):
        raise ValueError("Checkpoint dir '{}' is not a directory.".format(checkpoint_dir))

    runs = [f for f in os.listdir(checkpoint_dir) if not f.startswith('_') and not f.startswith('.')]
    if runs:
        run = int(sorted(runs, key=lambda x: int(x))[-1])
        if not resume:
            run += 1
    else:
        if resume:
            raise ValueError("Can't resume since no experiments were run before in checkpoint dir '{}'.".format(checkpoint_dir))
        run = 1
    experiment_folder = os.path.join(checkpoint_dir, str(run))
    if not resume:
        os.mkdir(experiment_folder)
    flag_path = os.path.join(experiment_folder, FLAG_FILE)
    resume_checkpoint = None

    _load_flags(model_config, data_config)
    flags = parse_flags()
    assert_all_flags_parsed()

    if resume:
        restored = json_load(flag_path)
        flags.update(restored)
        _restore_flags(flags)
        model_files = find_model_files(experiment_folder)
        if model_files:
            resume_checkpoint = model_files[max(model_files.keys())]
    else:
        try:
            flags['git_commit'] = get_git_revision_hash()
        except (subprocess.CalledProcessError, OSError):
            pass
        json_store(flag_path, flags)
        for src in (model_config, data_config):
            shutil.copy(src, os.path.join(experiment_folder, os.path.basename(src)))
        with open(os.path.join(experiment_folder, GIT_DIFF_FILE), 'a') as f:
            f.write(datetime.datetime.now().strftime('%c') + '\n')
    EXPERIMENT_FOLDER = experiment_folder
    return experiment_folder, resume_checkpoint


def load_from_checkpoint(checkpoint_dir, checkpoint_iter, path_prefix='', mode=None, **kwargs):
    """Rebuild data + model from a run folder (torch branch of the real function): flags.json -> flags, the stored copies of
    the config files -> loaders and model, `model.ckpt-<iter>` -> state dict."""
    from attrdict import AttrDict
    flags = AttrDict(json_load(os.path.join(checkpoint_dir, FLAG_FILE)))
    data_config = os.path.join(path_prefix, flags.data_config)
    model_config = os.path.join(path_prefix, flags.model_config)
    data = load(data_config, flags)
    model = load(model_config, flags)
    checkpoint_path = os.path.join(checkpoint_dir, 'model.ckpt-{}'.format(checkpoint_iter))
    if mode == 'torch':
        import torch
        checkpoint = torch.load(checkpoint_path, weights_only=False)
        model.load_state_dict(checkpoint['model_state_dict'])
    elif mode is None:
        checkpoint = checkpoint_path
    else:
        raise ValueError('Unkown mode: "{}".'.format(mode))
    return data, model, checkpoint


# ------------------------------------------------------------------------------------------------- printing
def fprint(text, timestamp=False, printonly=False):
    """Print `text` and append it to EXPERIMENT_FOLDER/FPRINT_FILE (reference experiment_tools.py:444-457)."""
    if timestamp:
        text = datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S') + ' ' + str(text)
    if EXPERIMENT_FOLDER is not None and not printonly:
        with open(os.path.join(EXPERIMENT_FOLDER, FPRINT_FILE), 'a+') as f:
            print(text, file=f)
    print(text)
