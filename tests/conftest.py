import os
import sys

# The single-GPU emulation of N ranks (tests/test_parity_gpu.py::test_fused_sharded_step_emulated_ranks) keeps a spinning barrier
# kernel of "rank 0" resident while the host enqueues the kernels of "rank 1".  With CUDA's default LAZY module loading the first
# launch of any kernel loads it at launch time, which synchronises the context — i.e. waits for the spinning kernel, which waits
# for work the blocked host thread has not enqueued yet: a deadlock that only the barrier's 2 s bound resolves (observed on B200:
# rank 0 then proceeds with incomplete peer data).  One process per GPU — the real deployment — cannot deadlock this way (a peer's
# host thread is never blocked by this process's loads).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
