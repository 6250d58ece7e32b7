import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libs():
    """Build the product .so and the oracle .so if they are missing (the GPU
    box receives prebuilt files with the snapshot)."""
    import subprocess
    lib = os.path.join(ROOT, "auditory_b200", "lib", "libauditory_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "auditory_b200", "csrc")])
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])


def pytest_sessionfinish(session, exitstatus):
    """Write the recorded parity margins (tests/util.py) when the GPU tests ran."""
    try:
        import json
        import util
        if not util.RECORDED:
            return
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as f:
            json.dump({"exitstatus": int(exitstatus), "errors": util.RECORDED}, f, indent=1, sort_keys=True)
    except Exception:
        pass
