"""The EXACT scenes bench.py times (BASELINE configs 2, 3, 4, 5), at full size, byte-compared with the CPU oracle.

bench.workload() builds the scene; the inputs are seeded CPU frames of the benchmark's resolution.  Reference scenes
these configurations restate: integration-tests/src/render_tests/tiles.rs:73-143, rescaler.rs:814-859."""
import numpy as np
import pytest

import smelter_b200 as s
import bench
from tests import harness
from tests.parity import OUTPUT_ID, assert_identical, nv12_frame, run_case

pytestmark = pytest.mark.gpu

NV12 = s.OutputFrameFormat.Nv12WgpuTexture


def bench_frames(wl, seed):
    """n distinct NV12 frames of the workload's input size: band-limited content with per-pixel noise on every third"""
    fr = {}
    for i in range(1, wl["n"] + 1):
        mk = harness.random_yuv420 if i % 3 == 0 else harness.smooth_yuv420
        fr[f"input_{i}"] = nv12_frame(mk(seed + i, wl["iw"], wl["ih"]), wl["iw"], wl["ih"])
    return fr


def check_workload(name, seed):
    wl = bench.workload(name)
    fr = bench_frames(wl, seed)
    got, exp, r = run_case(wl["scene"], fr, resolution=s.Resolution(wl["W"], wl["H"]), out_format=NV12, mode=wl["mode"])
    assert_identical(got, exp, f"bench scene {name}")
    return r


def test_cfg3_bench_scene_full_size():
    """16 x 4K NV12 -> 4K NV12, Tiles 4x4, Lanczos3 4:1, rounded tiles and the translucent rounded overlay"""
    r = check_workload("cfg3", 7000)
    kt = r.stats()
    assert kt["kernel_launches"] > 0


def test_cfg3b_bench_scene_full_size():
    """16 x 1080p -> 4K (2:1, 13 taps): the second integer-ratio kernel on the benchmark's own scene"""
    check_workload("cfg3b", 7100)


def test_cfg2_bench_scene_full_size():
    """4 x 1080p NV12 -> 1080p NV12, Tiles 2x2, CpuOptimized (bilinear in the composite, gamma blend)"""
    check_workload("cfg2", 7200)


def test_cfg4_bench_scenes_eight_outputs_one_tick():
    """8 outputs of one tick (one k_composite_multi launch, resamples shared between outputs), every output
    compared with the oracle"""
    wl = bench.workload("cfg4")
    fr = bench_frames(wl, 7300)
    r = s.Renderer(s.RendererOptions(rendering_mode=wl["mode"]))
    for iid in fr:
        r.register_input(iid)
    res = s.Resolution(wl["W"], wl["H"])
    scenes = {f"output_{k + 1}": bench.cfg4_scene(k, wl["n"]) for k in range(wl["n_out"])}
    for oid, sc in scenes.items():
        r.update_scene(oid, res, NV12, sc)
    out = r.render(s.FrameSet(frames=fr, pts=0.0))
    for oid, sc in scenes.items():
        used = {c.input_id for c in sc.children}
        _, exp, _ = run_case(sc, {k: v for k, v in fr.items() if k in used}, resolution=res, out_format=NV12, mode=wl["mode"])
        assert_identical([np.asarray(p) for p in out.frames[oid].data.planes], exp, f"cfg4 {oid}")


@pytest.mark.slow
def test_grid25_bench_scene_full_size():
    """25 x 4K -> 4K, Tiles 5x5 (5:1): box pre-decimation + Lanczos through the generic resampler passes
    (resampler.rs:56-67,305-340, downsample.wgsl:28-41) at full size"""
    check_workload("grid25", 7600)


@pytest.mark.slow
def test_cfg5_bench_scene_full_size():
    """32 x 4K -> 8K, 6x6 grid with margins (any-ratio kernel, fractional tile positions), radius + box shadows"""
    check_workload("cfg5", 7500)
