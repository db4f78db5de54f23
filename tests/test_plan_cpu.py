"""Host-side launch planner of the tcgen05 convolution kernel (spaa_conv_tc_plan: pure host arithmetic, no GPU): every ShadingNetSPAA layer at the
BASELINE shapes gets a plan that respects the hardware limits the kernel relies on."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plans(batch=None, hw=None, variant="shading"):
    spec = importlib.util.spec_from_file_location("plan_table", os.path.join(ROOT, "tools", "plan_table.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.layer_plans(batch, hw, variant)


@pytest.mark.parametrize("batch,hw,variant", [(32, (240, 320), "shading"), (24, (240, 320), "shading"), (1, (240, 320), "shading"), (100, (240, 320), "shading"),
                                              (32, (256, 256), "compen"), (8, (64, 96), "shading")])
def test_every_layer_has_a_plan_within_the_sm_limits(batch, hw, variant):
    rows, names = _plans(batch, hw, variant)
    assert len(rows) == 27
    for layer, p in rows:
        assert p is not None, f"{layer}: not covered by the halo kernel"
        d = dict(zip(names, p))
        assert d["ctas"] in (1, 2) and d["eg"] in (1, 2, 4) and d["nbuf"] in (1, 2, 4) and d["eg"] <= max(d["nbuf"], 1), (layer, d)
        assert d["threads"] == 64 + 128 * d["eg"] <= 640, (layer, d)
        # shared memory: <= 227 KB per CTA, and all resident CTAs of an SM (+1 KB each reserved by the driver) within 228 KB
        assert d["smem"] <= 227 * 1024 and d["ctas"] * (d["smem"] + 1024) <= 228 * 1024, (layer, d)
        assert 2 <= d["sa"] <= 8 and (d["resident"] == 1) == (d["sb"] == 0), (layer, d)
        assert d["S"] == 0 or 2 <= d["S"] <= 8, (layer, d)
        # registers: the narrow kernels are compiled for <= 96 registers (640-thread bound), the others for <= 168 at <= 320 threads
        assert d["ctas"] * d["threads"] * (96 if d["threads"] > 320 or d["ctas"] * d["threads"] > 384 else 168) <= 65536, (layer, d)


def test_plan_rejects_unsupported_descriptors():
    import ctypes
    from spaa_b200._lib import ConvDesc, SpaaError, lib
    d = ConvDesc()                      # all-zero descriptor: not a tensor-core shape
    p = (ctypes.c_int32 * 12)()
    with pytest.raises(SpaaError):
        lib().spaa_conv_tc_plan(ctypes.byref(d), 0, 0, 0, p)
