"""-m gpu: every kernel variant that stays in the tree for A/B runs is still correct.  The selectors are environment
variables read once per process, so each variant runs in its own subprocess (scripts/check_variant.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "default (producer/consumer forward, MN-major reverse)": {},
    "reverse: FP32 FFMA (nqmc_orbital.cu)": {"NQMC_TC_BWD": "0"},
    "value: layer 1 on the FP32 pipe (nqmc_orbital_val_tc.cu)": {"NQMC_VAL_L1TC": "0"},
    "everything FP32 FFMA": {"NQMC_TC": "0"},
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_matches_oracle(name):
    env = dict(os.environ)
    env.update(VARIANTS[name])
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "check_variant.py")], capture_output=True, text=True,
                         timeout=300, env=env, cwd=ROOT)
    assert "VARIANT_OK" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]
