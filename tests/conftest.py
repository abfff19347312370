import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu`")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_cfg():
    from oracle import mvldm_oracle as O
    return O.OracleCfg()


@pytest.fixture(scope="session")
def oracle_weights(oracle_cfg):
    from oracle import mvldm_oracle as O
    return O.init_weights(oracle_cfg, seed=0)


@pytest.fixture(scope="session")
def gpu_models(oracle_weights):
    """impl -> MultiViewUNet on cuda:0 loaded with the oracle's seed-0 weights (built lazily, cached)."""
    import mvldm_b200 as mv
    cache = {}

    def get(impl=0, graph=False):
        key = (impl, graph)
        if key not in cache:
            m = mv.MultiViewUNet(mv.default_cfg(), 11, 4, impl=impl, use_cuda_graph=graph)
            m.load_state_dict(oracle_weights)
            cache[key] = m.cuda().eval()
        return cache[key]

    yield get
    cache.clear()
