"""The drop-in boundary shown from the reference's side: the reference's OWN training entry point (train.py, unchanged) runs

  * on the CPU with the reference's models/vae_config.py and this repo's synthetic data config + Forge stand-in -- BASELINE
    config c1, "plumbing, no GPU" (SURVEY.md section 7 step 0);
  * on a B200 with THIS repo's plug-in files passed as --model_config (the boundary of SURVEY.md section 8b): fet.load imports
    the file by path, calls load(cfg), train.py builds its GECO loss from the returned losses, calls backward() and
    torch.optim.Adam.step(), saves / restores checkpoints, runs evaluation() with the ARI / segmentation-covering metrics and
    visualise_outputs() with model.sample().

The reference root is /root/reference where it exists (the build container) and otherwise oracle/_ref, the byte-identical copy
made by oracle/build_ref.py (git-ignored, shipped to the GPU box); MANIFEST.json proves the copy is unmodified."""
import hashlib
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import build_ref

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
RUNNER = os.path.join(HERE, 'train_py_runner.py')
DATA = os.path.join(REPO, 'genesis_b200', 'datasets', 'synthetic_config.py')
PLUGINS = os.path.join(REPO, 'genesis_b200', 'model_configs')

needs_ref = pytest.mark.skipif(build_ref.root() is None, reason='no reference checkout and no oracle/_ref copy (run python -m oracle.build_ref '
                                                                  'where /root/reference exists)')


def run_train(root, results, extra, timeout=900):
    cmd = [sys.executable, RUNNER, root, '--debug', '--data_config', DATA, '--results_dir', results, '--run_name', 'dropin'] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    return r, os.path.join(results, 'dropin')


def check_run_folder(rundir, n=1):
    d = os.path.join(rundir, str(n))
    names = set(os.listdir(d))
    assert {'flags.json', 'fprint.txt', 'model.ckpt-latest', 'model.ckpt-FINAL', 'synthetic_config.py'} <= names, names
    flags = json.load(open(os.path.join(d, 'flags.json')))
    for key in ('K_steps', 'img_size', 'batch_size', 'model_config', 'data_config', 'geco', 'debug'):
        assert key in flags, key
    log = open(os.path.join(d, 'fprint.txt')).read()
    assert 'FINAL VALIDATION STATS' in log and '[10/' in log
    ckpt = torch.load(os.path.join(d, 'model.ckpt-FINAL'), map_location='cpu', weights_only=False)
    assert {'model_state_dict', 'optimiser_state_dict', 'beta', 'iter_idx', 'err_ema'} <= set(ckpt)
    return flags, log, ckpt


def test_vendored_reference_is_unmodified():
    """oracle/_ref (when present) is a byte-identical copy: every file matches the sha256 recorded at copy time, and -- where the
    live checkout exists -- the checkout itself."""
    if not build_ref.available():
        pytest.skip('oracle/_ref not built')
    man = json.load(open(os.path.join(build_ref.DST, 'MANIFEST.json')))
    assert 'train.py' in man['files'] and 'models/genesis_config.py' in man['files']
    for rel, sha in man['files'].items():
        assert hashlib.sha256(open(os.path.join(build_ref.DST, rel), 'rb').read()).hexdigest() == sha, rel
        live = os.path.join(build_ref.SRC, rel)
        if os.path.exists(live):
            assert hashlib.sha256(open(live, 'rb').read()).hexdigest() == sha, rel


@needs_ref
def test_c1_reference_train_py_runs_unchanged_on_cpu(tmp_path):
    """BASELINE config c1: models/vae_config.py, 64x64 Multi-dSprites-shaped synthetic batches, batch 16, CPU -- the reference's
    train.py --debug (10 iterations, validation every 5, checkpoints, visualisation, final validation) through the Forge
    stand-in, then a --resume run that restores the flags and the newest checkpoint."""
    root = build_ref.root()
    model = os.path.join(root, 'models', 'vae_config.py')
    extra = ['--patch-vae-k-steps', '--gpu=False', '--model_config', model, '--batch_size', '16']
    r, rundir = run_train(root, str(tmp_path), extra)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    flags, log, ckpt = check_run_folder(rundir)
    assert flags['batch_size'] == 2 and flags['synthetic_kind'] == 'multid'       # --debug forces batch 2 (train.py:101-106)
    assert 'vae_config.py' in os.listdir(os.path.join(rundir, '1'))
    assert any(k.startswith('vae.q_z_nn') or 'q_z_nn' in k for k in ckpt['model_state_dict'])
    r2, _ = run_train(root, str(tmp_path), extra + ['--resume'])
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-3000:]
    assert 'Restoring checkpoint from' in r2.stdout and 'Starting training at iter = 11' in r2.stdout
    assert sorted(os.listdir(rundir)) == ['1']                                     # resumed in place, no new run folder


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('plugin,extra', [
    ('genesis_config.py', ['--K_steps', '5']),
    ('genesisv2_config.py', ['--K_steps', '4', '--synthetic_kind', 'stacks']),
    ('monet_config.py', ['--K_steps', '4']),
    ('vae_config.py', []),
])
def test_reference_train_py_drives_the_plugin_on_gpu(tmp_path, plugin, extra):
    """train.py (reference, unchanged) --debug with --model_config <this repo's plug-in>: 10 optimisation steps with the caller's
    Python GECO and torch.optim.Adam on the plug-in's parameters, 3 evaluation() passes incl. ARI / segmentation covering on the
    plug-in's log_m_k, checkpoints in the reference's format, visualise_outputs() incl. model.sample().  The run must finish,
    the ELBO must be finite and the parameters must have moved."""
    root = build_ref.root()
    model = os.path.join(PLUGINS, plugin)
    r, rundir = run_train(root, str(tmp_path), ['--model_config', model] + extra)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'Use GPU: True' in r.stdout
    flags, log, ckpt = check_run_folder(rundir)
    assert plugin in os.listdir(os.path.join(rundir, '1'))
    elbos = [float(line.split('elb:')[1].split()[0]) for line in log.splitlines() if 'elb:' in line]
    assert len(elbos) >= 11 and all(e == e and abs(e) < 1e8 for e in elbos), elbos
    first = torch.load(os.path.join(rundir, '1', 'model.ckpt-0'), map_location='cpu', weights_only=False)['model_state_dict']
    moved = [k for k, v in ckpt['model_state_dict'].items() if v.is_floating_point() and not torch.equal(v, first[k])]
    assert len(moved) > 0.8 * sum(v.is_floating_point() for v in first.values()), len(moved)
    if plugin != 'vae_config.py':
        assert "'ari'" in log or 'ari' in log                                       # evaluation() computed the segmentation metrics
