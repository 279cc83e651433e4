import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_pkg():
    """The package directory is `go-vectorsearch_b200` (hyphen): import it under a legal module name."""
    name = "go_vectorsearch_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "go-vectorsearch_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(path, "__init__.py"),
                                                  submodule_search_locations=[path])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def vs(pkg):
    """The package with the CUDA backend initialised; fails loudly when there is no B200."""
    pkg._lib.init()
    return pkg


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.binding.build()
    return o
