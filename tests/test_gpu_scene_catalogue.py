"""The reference's whole in-scope render-test catalogue (tests/golden/ref_scenes.py: simple / view / rescaler / tiles /
transition / tiles_transitions, 111 tests, every snapshot pts) replayed on the GPU through the C ABI and byte-compared
with the CPU oracle; the oracle takes its layouts from the independent layout engine (tests/layout_ref.py), and the
product's flattened layouts are asserted identical to those on the way (tests/parity.py)."""
import numpy as np
import pytest

import smelter_b200 as s
from tests import harness
from tests import ref_scene_rt as rt
from tests.golden import ref_scenes
from tests.parity import OUTPUT_ID, TrackedRenderer, assert_identical, oracle_output, product_planes, yuv_frame

pytestmark = pytest.mark.gpu

CASES = [(m, n) for m, tests in ref_scenes.MODULES.items() for n in tests]
YUV = s.OutputFrameFormat.PlanarYuv420Bytes


def frame_of(inp, pts):
    w, h = inp.resolution.width, inp.resolution.height
    planes = harness.multiscale_grid(w, h) if inp.pattern == "multiscale_grid" else harness.test_input(inp.index, w, h)
    return yuv_frame(planes, w, h, pts)


@pytest.mark.parametrize("module,name", CASES)
def test_reference_scene(module, name):
    rec = rt.record(ref_scenes.MODULES[module][name])
    if rec is None:
        pytest.skip("text / image / shader components are outside the compositor hot path")
    r = TrackedRenderer(s.RendererOptions(rendering_mode=rec.mode))
    for i in rec.inputs:
        r.register_input(i.name)
    scene, n_snap = None, 0
    for kind, arg in rec.steps:
        if kind == "update":
            scene = arg
            r.update_scene(OUTPUT_ID, rec.resolution, YUV, scene)
            continue
        pts = arg
        frames = {i.name: frame_of(i, pts) for i in rec.inputs}
        out = r.render(s.FrameSet(frames=dict(frames), pts=pts))
        got = product_planes(out.frames[OUTPUT_ID])
        exp = oracle_output(r, scene, frames, rec.resolution, YUV, rec.mode, pts)
        assert_identical(got, exp, f"{module}/{name} pts={pts}")
        n_snap += 1
    assert n_snap > 0
