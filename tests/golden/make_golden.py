#!/usr/bin/env python
"""Regenerates tests/golden/oracle_golden.json: SHA-256 of the oracle's output planes for every case of cases.py.
Runs without a GPU (device-free library handle for the scene / layout maths + the CPU oracle).
    python -m tests.golden.make_golden
The reference itself cannot be executed here (Rust + wgpu), so these vectors pin the ORACLE (already pinned by the
reference's in-tree known-answer vectors, tests/test_oracle_kat.py) and every later change of oracle or kernels against
today's behaviour -- they are not outputs of the reference."""
import json
import os

import smelter_b200 as s
from tests.golden.cases import CASES, digest
from tests.parity import OUTPUT_ID, TrackedRenderer, oracle_output


def expected_planes(name):
    scene_f, frames_f, res, fmt, mode, pts = CASES[name]
    scene, frames = scene_f(), frames_f()
    r = TrackedRenderer(s.RendererOptions(rendering_mode=mode, cuda_device=-1))   # layouts from the independent engine
    for iid in frames:
        r.register_input(iid)
    r.update_scene(OUTPUT_ID, res, fmt, scene)
    r.debug_set_inputs(pts, {k: f.resolution for k, f in frames.items()})
    return oracle_output(r, scene, frames, res, fmt, mode, pts)


def main():
    out = {}
    for name in CASES:
        planes = expected_planes(name)
        out[name] = {"sha256": digest(planes), "shapes": [list(p.shape) for p in planes]}
        print(name, out[name]["sha256"][:16], out[name]["shapes"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
