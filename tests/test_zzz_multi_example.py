"""examples/example_multi.c on the GPU box: fsb200_lr_multi() on one GPU and on every visible GPU from plain C, results
compared bit for bit (sorted last on purpose: it is the newest test of the round)."""
import os
import subprocess

import pytest

from tests.test_c_example import test_multi_gpu_example_compiles_and_links as _compile

pytestmark = pytest.mark.gpu


def test_multi_gpu_example_runs(tmp_path):
    _compile(tmp_path)
    out = subprocess.run([os.path.join(tmp_path, "example_multi"), "60000"], check=True, capture_output=True, text=True).stdout
    assert "bit-identical to the one-GPU call" in out
