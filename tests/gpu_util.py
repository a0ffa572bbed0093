import pytest


def gpu_context():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import libmems_b200 as mems
    return mems.Context(0)
