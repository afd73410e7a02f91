"""Host-side pieces of the bulk generator (medfusion_b200/sample_dataset.py) — no GPU needed."""
import numpy as np
import pytest
import torch

from util import load_golden  # noqa: F401  (path setup)
from medfusion_b200.sample_dataset import chunks, generate_dataset
from util import to_uint8_hwc


def test_chunks_like_the_reference_script():
    # scripts/helpers/sample_dataset.py:10-13 with n_samples=7869, sample_batch=200 -> 39 full chunks + a tail of 69
    sizes = [len(c) for c in chunks(list(range(7869)), 200)]
    assert sizes == [200] * 39 + [69] and sum(sizes) == 7869
    assert list(chunks([], 3)) == []


def test_host_uint8_conversion_is_the_scripts_numpy_formula():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 3, 16, 16, generator=g) * 1.5          # values beyond [-1, 1] exercise the clip
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.999999])
    ref = x.numpy()
    ref = np.stack([np.moveaxis((im.clip(-1, 1) + 1) / 2 * 255, 0, -1).astype(np.uint8) for im in ref])   # :47-50
    assert np.array_equal(to_uint8_hwc(x).numpy(), ref)


def test_generate_dataset_refuses_cpu_pipelines():
    class P:
        device = torch.device("cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        generate_dataset(P(), 4, sink=lambda c, a: None)
