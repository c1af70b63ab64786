"""pytest configuration: `gpu` marker, repo paths, shared fixtures."""
from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_artefacts():
    """The shared library and the host CLI are build products (git-ignored): build them when a fresh checkout
    runs the tests before `__graft_entry__.build()` (nvcc cross-compiles without a GPU)."""
    lib = ROOT / "shafa-cd_b200" / "libshafa_cuda.so"
    cli = ROOT / "shafa-cd_b200" / "shafa"
    if not (lib.exists() and cli.exists()):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package

    return load_package()


@pytest.fixture(scope="session")
def orc():
    """The oracle (test infrastructure): CPU restatement of the reference."""
    from oracle import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def shafa(pkg):
    """GPU context through the C-ABI; fails loudly if the CUDA library is missing."""
    sh = pkg.Shafa(device=0)
    yield sh
    sh.close()


def small_cases():
    return sorted(p.name for p in GOLDEN.iterdir() if p.is_dir() and (p / "input.bin").exists())


def encode_only_cases():
    """Reference fixtures whose input has a one-symbol block: F, T, C only (the reference decoder crashes on them)."""
    d = GOLDEN / "encode_only"
    return sorted(p.name for p in d.iterdir() if (p / "input.bin").exists()) if d.is_dir() else []
