import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu_codec() -> bool:
    try:
        import slimfastq_b200 as S

        S.Codec().close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device (or without the built library) skips the gpu-marked tests instead of
    failing them; `-m gpu` on the GPU box still fails loudly if the library cannot create a context there."""
    if config.getoption("-m") and "gpu" in config.getoption("-m") and "not gpu" not in config.getoption("-m"):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not _have_gpu_codec():
        skip = pytest.mark.skip(reason="no CUDA device / libsfq_b200.so: the product path has no CPU fallback")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def codec():
    import slimfastq_b200 as S

    c = S.Codec()          # raises (no CPU fallback) when there is no GPU / no built library
    yield c
    c.close()


def sample_files():
    from oracle import oracle as O

    d = O.REF_SAMPLES
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".fq"))
