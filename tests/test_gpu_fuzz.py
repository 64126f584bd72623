"""GPU: randomised differential test of the three scoring paths (tools/fuzz_paths.py): random shapes, metric sets, flags, exact
ties, zero-factor users, noise on and off -- FMA tiles, tensor-core filter (+ hand-backs) and full-order path must agree bit for
bit.  A 25-second budget here (a few hundred cases); profiles/r02fin_fuzz_paths.txt holds a 200-second run (6161 cases)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_cases_agree_across_paths(rb, seed):
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_paths.py"), "25", str(seed)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "all paths identical" in run.stdout, run.stdout[-2000:] + run.stderr[-2000:]
