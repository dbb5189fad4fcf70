"""Run the reference's training entry point UNCHANGED (train.py:592-594: `main_flags(); main()`) from a reference root
(/root/reference, or the byte-identical copy under oracle/_ref made by oracle/build_ref.py) with this repo's stand-ins for
Forge / attrdict / simplejson / tensorboardX / tensorflow / imageio on sys.path.

    python tests/train_py_runner.py <reference_root> [--patch-vae-k-steps] <train.py flags ...>

--patch-vae-k-steps works around a bug of the reference itself: train.py:463 reads `model.K_steps`, which the reference's
BaselineVAE (models/vae_config.py:40-101) does not define, so `train.py --model_config models/vae_config.py` dies in
visualise_outputs at iteration 0 with an AttributeError.  The flag pre-imports that config file under the module name Forge
will look up (sys.modules hit, experiment_tools.py:269-271) and sets `BaselineVAE.K_steps = 1`; no file is modified."""
import os
import runpy
import sys


def main():
    root = os.path.abspath(sys.argv[1])
    args = sys.argv[2:]
    patch_vae = '--patch-vae-k-steps' in args
    args = [a for a in args if a != '--patch-vae-k-steps']
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(repo, 'genesis_b200', 'compat'), repo]
    os.chdir(root)
    sys.argv = [os.path.join(root, 'train.py')] + args
    if patch_vae:
        import forge.experiment_tools as fet
        path = [a.split('=', 1)[1] if '=' in a else args[i + 1] for i, a in enumerate(args) if a.startswith('--model_config')][0]
        module, _ = fet._import_module(path)
        module.BaselineVAE.K_steps = 1
    runpy.run_path(os.path.join(root, 'train.py'), run_name='__main__')


if __name__ == '__main__':
    main()
