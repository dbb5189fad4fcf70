"""CPU, opt-in (G2_RUN_EMU_MODELS=1; minutes per case): one training step of a plug-in with every kernel call routed to the CPU
emulation of the kernel sources, against the oracle (tests/cuda_emu/run_engine_emu.py).  Outputs of the round-1 runs of every
model, variant and experimental switch are kept in profiles/r01_emulated_engine_runs.txt."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.skipif(os.environ.get('G2_RUN_EMU_MODELS') != '1', reason='slow emulation runs (set G2_RUN_EMU_MODELS=1)')

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [
    ['genesis', '2', '2'],
    ['genesis', '2', '2', '--fused-latent'],
    ['genesisv2', '3', '2'],
    ['monet', '2', '2'],
    ['vae', '1', '2'],
    ['genesis', '2', '2', 'two_stage=False'],
    ['genesis', '2', '2', 'comp_symmetric=True'],
    ['genesisv2', '3', '2', 'klm_loss=True', 'detach_mr_in_klm=False'],
    ['genesisv2', '3', '2', 'kernel=laplacian'],
    ['genesisv2', '6', '3', 'dynamic_K=True', '--param-add', 'att_process.log_sigma=2.7726'],
    ['monet', '3', '2', 'prior_mode=scope'],
]


@pytest.mark.parametrize('case', CASES, ids=[' '.join(c) for c in CASES])
def test_engine_step_under_emulation(case):
    r = subprocess.run([sys.executable, os.path.join(HERE, 'cuda_emu', 'run_engine_emu.py')] + case, capture_output=True, text=True,
                       timeout=3000)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
