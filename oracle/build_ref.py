"""ORACLE / test infrastructure -- recipe that makes the REAL reference runnable where /root/reference does not exist (the GPU box).

    python -m oracle.build_ref            (also called by __graft_entry__.build() when /root/reference is present)

The reference is pure Python with no build step, so "building" it = copying the files its training entry point imports, byte for
byte, from where they lie under /root/reference into oracle/_ref/ (git-ignored: reference sources never enter the history;
NOT gpurun-ignored: the directory travels to the GPU box like a built .so).  Nothing is modified; MANIFEST.json records the
sha256 of every copied file so a test can prove that.  Users of oracle/_ref:

  * bench.py --impl reference / cpu_baseline   -- the reference's own nn.Modules (models/*_config.py) on the host cores
                                                  (`kind: "reference"`); falls back to the oracle port when _ref is absent
  * tests/test_train_py_dropin.py              -- the reference's unchanged train.py driving OUR plug-in (GPU) and its own
                                                  vae_config on the CPU (BASELINE config c1)

Third-party imports the reference needs and this image lacks (tensorflow, attrdict, simplejson, tensorboardX, imageio, Forge,
which itself needs TensorFlow) are satisfied by the stand-ins in genesis_b200/compat/."""
import hashlib
import json
import os
import shutil
import sys

SRC = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')

FILES = ['train.py']
TREES = ['models', 'modules', 'utils', 'scripts', 'third_party/sylvester', 'third_party/pytorch_fid']


def _sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def build(src=SRC, dst=DST, verbose=False):
    """Copy the reference's training path into dst; returns dst, or None when the reference checkout is not available."""
    if not os.path.isdir(src):
        return None
    manifest = {}
    os.makedirs(dst, exist_ok=True)
    todo = [(f, f) for f in FILES]
    for tree in TREES:
        for dirpath, _, files in os.walk(os.path.join(src, tree)):
            for f in files:
                if f.endswith(('.pyc', '.txt', '.md')) or f == 'LICENSE':
                    continue
                rel = os.path.relpath(os.path.join(dirpath, f), src)
                todo.append((rel, rel))
    for rel, out in todo:
        s, d = os.path.join(src, rel), os.path.join(dst, out)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or _sha(d) != _sha(s):
            shutil.copyfile(s, d)
        manifest[out] = _sha(d)
    json.dump({'source': src, 'files': manifest}, open(os.path.join(dst, 'MANIFEST.json'), 'w'), indent=1, sort_keys=True)
    if verbose:
        print('oracle/_ref: %d files from %s' % (len(manifest), src))
    return dst


def available(dst=DST):
    return os.path.exists(os.path.join(dst, 'MANIFEST.json'))


def root():
    """Directory to put on sys.path (and to chdir into: utils/misc.py:88 opens 'utils/colour_palette15.json' relative to
    the working directory): the live checkout when present, else the vendored copy, else None."""
    if os.path.isdir(SRC):
        return SRC
    return DST if available() else None


if __name__ == '__main__':
    out = build(verbose=True)
    sys.exit(0 if out else 1)
