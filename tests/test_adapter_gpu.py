"""-m gpu: the voxblox-side adapter (include/voxblox_b200/gpu_integrators.h), compiled against the
reference's own headers and run next to the reference's CPU integrators (oracle/adapter_test.cc).
The binary is built where /root/reference exists and travels with the repository snapshot."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "adapter_test")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/adapter_test not built (no /root/reference here)")
def test_drop_in_adapter_against_reference_classes():
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ADAPTER TEST OK" in out.stdout
