"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b200reg.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest

from conftest import PKG_NAME, ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "b200reg.h")).read()
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    builder = importlib.import_module(PKG_NAME + ".build")
    so = builder.build()
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 52
    for n in names:
        assert hasattr(lib, n), n
    binding = importlib.import_module(PKG_NAME + ".binding")
    assert sorted(binding.EXPORTS) == names
    assert binding.lib().b200_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    binding = importlib.import_module(PKG_NAME + ".binding")
    with pytest.raises(binding.B200Error) as e:
        binding.Context(0)
    assert e.value.code == binding.ERR_NODEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_oracle():
    """Nothing in the package, include/ or apps/ may import, link or name the oracle."""
    for base in (os.path.join(ROOT, PKG_NAME), os.path.join(ROOT, "include"), os.path.join(ROOT, "apps")):
        for dp, _, files in os.walk(base):
            if "build" in dp.split(os.sep) or "bin" in dp.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "pcl_oracle" not in txt and "orc_" not in txt, os.path.join(dp, f)


def test_keypoint_extractors(synth):
    m = synth.make_model("y", 4000)
    vg = synth.voxel_grid(m, 0.02)
    us = synth.uniform_sampling(m, 0.02)
    assert len(vg) == len(us) and 100 < len(vg) < 2000
    # uniform sampling returns input points; voxel grid returns centroids inside the voxel
    assert all((m == u).all(1).any() for u in us[:20])
    assert np.all(np.floor(vg / np.float32(0.02)) == np.floor(us / np.float32(0.02)))


def test_partial_view_text_format(tmp_path):
    """The reference dumps a view's descriptors as text, one float per line (CAD_desc.cpp:354-370); the reader and
    writer of that format need no GPU."""
    import importlib
    import numpy as np
    from conftest import PKG_NAME
    binding = importlib.import_module(PKG_NAME + ".binding")
    rng = np.random.Generator(np.random.PCG64(1))
    d = rng.uniform(0, 0.3, (7, 352)).astype(np.float32)
    d[2, 5] = 0.0
    p = str(tmp_path / "Partial_View3.txt")
    binding.write_partial_view_text(p, d)
    lines = open(p).read().split("\n")
    assert len(lines) == 7 * 352 + 1 and lines[2 * 352 + 5] == "0"
    back = binding.read_partial_view_text(p)
    assert back.shape == (7, 352) and np.abs(back - d).max() <= 5e-7 * 3 + 1e-6 * np.abs(d).max()
    open(p, "a").write("0.5\n")
    import pytest
    with pytest.raises(ValueError):
        binding.read_partial_view_text(p)
