"""The training loop end to end on the GPU (scripts/train_smoke.py): a 64^2 split-FC config trains through the captured
CUDA graphs, writes reference-format checkpoints, resumes (exact optimiser step counts) and the run directory loads
through the inference front end."""
import os
import runpy

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_smoke(capsys):
    runpy.run_path(os.path.join(ROOT, 'scripts', 'train_smoke.py'), run_name='__main__')
    assert 'train smoke ok' in capsys.readouterr().out
