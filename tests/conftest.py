import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """-> (oracle ScpInstance, npz dict) from tests/golden/<name>.npz"""
    from oracle import scp_io
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    m, n0 = int(z["m"]), int(z["n_orig"])
    offs = z["row_offs"].astype(np.int64)
    cols = z["col_inds"].astype(np.int32)
    rows = [cols[offs[i]:offs[i + 1]] for i in range(m)]
    inst = scp_io.to_standard_form(m, n0, z["costs"].astype(np.float64), rows, name)
    return inst, z


def node_from_instance(inst, **env_kw):
    import sypha_b200 as sb
    env = sb.SyphaEnvironment(**env_kw)
    return sb.SyphaNodeSparse.from_csr(inst.m, inst.n, inst.n_orig, inst.offs, inst.inds, inst.vals,
                                       inst.c, inst.b, env)


@pytest.fixture(scope="session")
def cuda_ws():
    """One persistent workspace for the whole GPU session (like the B&B driver's IpmWorkspace)."""
    import sypha_b200 as sb
    ws = sb.IpmWorkspace()
    sb.initializeIpmWorkspace(ws)
    yield ws
    sb.releaseIpmWorkspace(ws)
