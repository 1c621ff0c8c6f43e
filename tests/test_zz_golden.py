"""Committed golden digests (tests/golden/oracle_golden.json, made by tests/golden/make_golden.py):
* CPU: the oracle still produces exactly those bytes (guards the checker itself against drift);
* GPU: the CUDA path produces exactly those bytes, without consulting the live oracle."""
import json
import os

import numpy as np
import pytest

import smelter_b200 as s
from tests.golden.cases import CASES, digest
from tests.golden.make_golden import expected_planes
from tests.parity import OUTPUT_ID

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")))


def test_golden_file_covers_every_case():
    assert sorted(GOLDEN) == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_committed_golden(name):
    planes = expected_planes(name)
    assert [list(p.shape) for p in planes] == GOLDEN[name]["shapes"]
    assert digest(planes) == GOLDEN[name]["sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_committed_golden(name):
    scene_f, frames_f, res, fmt, mode, pts = CASES[name]
    scene, frames = scene_f(), frames_f()
    r = s.Renderer(s.RendererOptions(rendering_mode=mode))
    for iid in frames:
        r.register_input(iid)
    r.update_scene(OUTPUT_ID, res, fmt, scene)
    out = r.render(s.FrameSet(frames=dict(frames), pts=pts)).frames[OUTPUT_ID].data.planes
    planes = [np.asarray(p).reshape(shape) for p, shape in zip(out, GOLDEN[name]["shapes"])]
    assert len(planes) == len(GOLDEN[name]["shapes"])
    assert digest(planes) == GOLDEN[name]["sha256"], f"{name}: CUDA output differs from the committed golden bytes"
