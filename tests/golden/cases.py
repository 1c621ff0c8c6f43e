"""Scene-level golden cases: scenes re-typed from the reference's render tests with seeded / procedural inputs.
`make_golden.py` computes the expected output planes with the CPU oracle (fed with the layouts the host library
flattens, on a device-free handle) and commits their SHA-256 digests; `tests/test_golden.py` checks both the oracle
(regression, CPU) and the CUDA path (GPU) against those committed digests.  Test infrastructure."""
import numpy as np

import smelter_b200 as s
from tests import harness
from tests.parity import nv12_frame, wide_chroma_frame, yuv_frame

BG = s.RGBAColor(51, 51, 51, 255)   # #333333FF, tiles.rs:85-95
V = s.ViewComponent
YUV, NV12, RGBA = (s.OutputFrameFormat.PlanarYuv420Bytes, s.OutputFrameFormat.Nv12WgpuTexture,
                   s.OutputFrameFormat.RgbaWgpuTexture)
GPU, CPU = s.RenderingMode.GpuOptimized, s.RenderingMode.CpuOptimized


def streams(n):
    return [s.InputStreamComponent(input_id=f"input_{i}") for i in range(1, n + 1)]


def inputs(n, w=640, h=360):
    return {f"input_{i}": yuv_frame(harness.test_input(i, w, h), w, h) for i in range(1, n + 1)}


def _overlay_scene():
    kids = [s.RescalerComponent(child=c, border_radius=s.BorderRadius.new_with_radius(12.0)) for c in streams(4)]
    over = V(position=s.Position.Absolute(width=300.0, height=80.0, left=170.0, bottom=20.0),
             background_color=s.RGBAColor(16, 32, 160, 112), border_radius=s.BorderRadius.new_with_radius(16.0))
    return V(background_color=BG, children=[s.TilesComponent(children=kids, background_color=BG), over])


def _view_scene():
    sh = [s.BoxShadow(offset_x=12.0, offset_y=18.0, blur_radius=20.0, color=s.RGBAColor(0, 0, 0, 200))]
    inner = V(children=streams(1), position=s.Position.Static(width=320.0, height=180.0),
              border_radius=s.BorderRadius(40.0, 8.0, 24.0, 60.0), border_width=6.0,
              border_color=s.RGBAColor(255, 255, 0, 255), box_shadow=sh)
    outer = V(children=[inner], position=s.Position.Absolute(width=400.0, height=260.0, left=120.0, top=50.0),
              border_radius=s.BorderRadius(80.0, 20.0, 50.0, 10.0), padding=s.Padding(10, 10, 10, 120),
              background_color=s.RGBAColor(0, 0, 255, 128))
    return V(children=[outer], background_color=BG)


def _cfg3_third():
    return {f"input_{i}": nv12_frame(harness.random_yuv420(300 + i, 1280, 720) if i % 2 else
                                     harness.smooth_yuv420(300 + i, 1280, 720), 1280, 720) for i in range(1, 5)}


# name -> (scene, frames, resolution, output format, mode, pts)
CASES = {
    "tiles_02_inputs_yuv420": (lambda: s.TilesComponent(children=streams(2), background_color=BG), lambda: inputs(2),
                               s.Resolution(640, 360), YUV, GPU, 0.0),
    "tiles_05_inputs_nv12": (lambda: s.TilesComponent(children=streams(5), background_color=BG), lambda: inputs(5),
                             s.Resolution(640, 360), NV12, GPU, 0.0),
    "tiles_03_inputs_rgba": (lambda: s.TilesComponent(children=streams(3), background_color=BG), lambda: inputs(3),
                             s.Resolution(640, 360), RGBA, GPU, 0.0),
    "config3_third_size_4to1_overlay": (_overlay_scene, _cfg3_third, s.Resolution(640, 360), NV12, GPU, 0.0),
    "view_nested_radius_border_shadow": (_view_scene, lambda: inputs(1), s.Resolution(640, 360), YUV, GPU, 0.0),
    "tiles_02_cpu_optimized_bilinear": (lambda: s.TilesComponent(children=streams(2), background_color=BG),
                                        lambda: inputs(2), s.Resolution(640, 360), YUV, CPU, 0.0),
    "pass_through_rescaled": (lambda: s.InputStreamComponent(input_id="input_1"), lambda: inputs(1),
                              s.Resolution(400, 300), YUV, GPU, 0.0),
    "uyvy_and_yuv444_tiles_422_out": (lambda: s.TilesComponent(children=streams(2), background_color=BG),
                                      lambda: {"input_1": wide_chroma_frame("InterleavedUyvy422", 11, 640, 360),
                                               "input_2": wide_chroma_frame("PlanarYuv444", 12, 640, 360)},
                                      s.Resolution(640, 360), s.OutputFrameFormat.PlanarYuv422Bytes, GPU, 0.0),
}


def digest(planes):
    import hashlib
    h = hashlib.sha256()
    for p in planes:
        a = np.ascontiguousarray(np.asarray(p), dtype=np.uint8)
        h.update(repr(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()
