"""Test configuration: puts the package directory and the oracle on sys.path and registers the gpu marker.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI load/export checks.
``-m gpu`` runs on a B200: the parity tests proper, every one calling through the C ABI.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "pro-b-gan_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE_SCRIPT = Path("/root/reference/pro_b_gan_infer.py")
# the copy __graft_entry__.build() leaves for the GPU box (the reference mount does not exist there)
REFERENCE_SCRIPT_COPY = ROOT / "oracle" / "_ref" / "pro_b_gan_infer.py"
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    from oracle import prot_b_gan_oracle
    return prot_b_gan_oracle


@pytest.fixture(scope="session")
def synth():
    from pbg import synth as s
    return s


@pytest.fixture(scope="session")
def oracle_models(oracle, synth):
    """(G, D) oracle modules on CPU with the frozen seeds."""
    return synth.make_models(oracle.ModularGenerator, oracle.ModularDiscriminator)


@pytest.fixture(scope="session")
def tables(synth):
    return synth.make_tables()
