import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, 'multimodal-fusion-fpn_b200')
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box via gpurun)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def mirror():
    """Import the reference-shaped host mirror (config.py parses sys.argv at import, like the reference)."""
    import contextlib
    import io
    saved = sys.argv
    sys.argv = ['x', '--training-dataset', 'hrf_fusion', '--model', 'FPNHybridFusion', '--fusion-modality', 'slo',
                '--crop', 'relative_2d_max']
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import config as cfg
            from models import fusion_nets
            from common import loss, weight_init, pl_model_wrapper
    finally:
        sys.argv = saved

    class M:
        pass
    m = M()
    m.config, m.fusion_nets, m.loss, m.weight_init, m.wrapper = cfg.config, fusion_nets, loss, weight_init, pl_model_wrapper

    def build(name='FPNHybridFusion', crop='relative_2d_max', modality='slo', n_out=1):
        cfg.config.crop, cfg.config.fusion_modality, cfg.config.number_of_outputs = crop, modality, n_out
        with contextlib.redirect_stdout(io.StringIO()):
            return fusion_nets.factory_classes[name]()
    m.build = build
    return m
