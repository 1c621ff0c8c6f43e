"""Text nodes: TextRendererNode::render (smelter-render/src/transformations/text_renderer.rs:72-167) =
clear to the background colour + glyphon's prepared glyph quads alpha-blended from the mask / colour atlas.

CPU part: the oracle twin (orc_render_text) against hand-derived values and against an independent per-pixel restatement
in Python (exact rational fma, tables re-derived from the sRGB formulas).  GPU part: the product (smr_render_text ->
k_text) byte-exact against the oracle, and a text texture used as a layer of a scene.
glyphon is an un-vendored git dependency of the reference (0.11.0 @ smelter-labs c784922): parity unpinned beyond its
published shader / blend state, which is what both sides restate.
"""
from fractions import Fraction

import numpy as np
import pytest

from oracle import oracle as orc

GD = np.dtype(orc.GLYPH_DTYPE)
COLOR, MASK = 0, 1


def soft_mask_atlas(w, h, seed):
    """an R8 atlas of anti-aliased blobs: every coverage value occurs, with exact 0 and exact 255 regions"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    cov = np.clip(1.4 * (np.sin(xx / 3.1 + seed) * np.cos(yy / 2.3) + 0.3), 0.0, 1.0)
    a = np.rint(cov * 255).astype(np.uint8)
    a[:2, :] = np.arange(w, dtype=np.uint8)[None, :] if w <= 256 else a[:2, :]
    a[rng.random((h, w)) < 0.03] = 255
    return a


def color_atlas(w, h, seed):
    """an RGBA8 colour atlas (emoji-like): premultiplied-looking texels with every alpha"""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    t[..., 3] = (np.arange(w)[None, :] * 7 + np.arange(h)[:, None] * 3) % 256
    t[::5, ::3, 3] = 255
    t[1::7, ::4, 3] = 0
    return t


def random_glyphs(n, w, h, aw, ah, seed, contents=(MASK, COLOR)):
    rng = np.random.default_rng(seed)
    g = np.zeros(n, GD)
    g["width"] = rng.integers(0, 24, n)
    g["height"] = rng.integers(0, 20, n)
    g["x"] = rng.integers(-12, w + 4, n)
    g["y"] = rng.integers(-10, h + 4, n)
    g["atlas_x"] = rng.integers(0, max(aw - 24, 1), n)
    g["atlas_y"] = rng.integers(0, max(ah - 20, 1), n)
    g["color"] = rng.integers(0, 256, (n, 4))
    g["color"][::3, 3] = 255
    g["content"] = rng.choice(contents, n)
    return g


# ---- independent restatement -------------------------------------------------------------------------------------
def _f32_of(q):
    """the float32 nearest to the rational q, ties to even (no double rounding)"""
    c = np.float32(float(q))
    cands = {float(c), float(np.nextafter(c, np.float32(-np.inf))), float(np.nextafter(c, np.float32(np.inf)))}
    best = None
    for v in cands:
        d = abs(Fraction(v) - q)
        even = (np.float32(v).view(np.uint32) & 1) == 0
        if best is None or d < best[0] or (d == best[0] and even):
            best = (d, v)
    return np.float32(best[1])


def _fma(a, b, c):
    return _f32_of(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def _eotf(x):
    return x / 12.92 if x <= 0.04045 else ((x + 0.055) / 1.055) ** 2.4


_DEC = [np.float32(_eotf(i / 255.0)) for i in range(256)]                    # NC-3
_THR = [np.float32(_eotf((k + 0.5) / 255.0)) for k in range(255)]            # NC-4 decision thresholds
_U8N = [np.float32(i) / np.float32(255.0) for i in range(256)]               # NC-1


def _enc(x):
    x = min(max(np.float32(x), np.float32(0)), np.float32(1))
    return int(np.searchsorted(np.array(_THR, np.float32), x, side="right"))


def _u8(x):
    x = min(max(np.float32(x), np.float32(0)), np.float32(1))
    return int(np.rint(np.float32(x * np.float32(255.0))))


def text_ref(w, h, bg, glyphs, mask, color, color_mode, mode):
    a = bg[3] / 255.0
    lin = (lambda c: _eotf(c / 255.0) if c / 255.0 >= 0.04045 else (c / 255.0) / 12.92) if mode == 0 else (lambda c: c / 255.0)
    clear = [np.float32(a * lin(bg[c])) for c in range(3)] + [np.float32(a)]
    store = _enc if mode == 0 else _u8
    dlut = _DEC if mode == 0 else _U8N
    clut = _DEC if color_mode == 0 else _U8N
    out = np.zeros((h, w, 4), np.uint8)
    out[:, :] = [store(clear[0]), store(clear[1]), store(clear[2]), _u8(clear[3])]
    for G in glyphs:
        for dy in range(int(G["height"])):
            for dx in range(int(G["width"])):
                px, py = int(G["x"]) + dx, int(G["y"]) + dy
                if not (0 <= px < w and 0 <= py < h):
                    continue
                ax, ay = int(G["atlas_x"]) + dx, int(G["atlas_y"]) + dy
                if G["content"] == MASK:
                    t = mask[min(ay, mask.shape[0] - 1), min(ax, mask.shape[1] - 1)]
                    s = [clut[G["color"][0]], clut[G["color"][1]], clut[G["color"][2]], np.float32(_U8N[G["color"][3]] * _U8N[t])]
                else:
                    t = color[min(ay, color.shape[0] - 1), min(ax, color.shape[1] - 1)]
                    s = [clut[t[0]], clut[t[1]], clut[t[2]], _U8N[t[3]]]
                sa = s[3]
                ia = np.float32(np.float32(1.0) - sa)
                d = out[py, px]
                for c in range(3):
                    d[c] = store(_fma(dlut[d[c]], ia, np.float32(s[c] * sa)))
                d[3] = _u8(_fma(_U8N[d[3]], ia, sa))
    return out


# ---- CPU: the oracle twin ------------------------------------------------------------------------------------------
def test_glyph_record_is_24_bytes():
    from smelter_b200 import _ffi as F
    assert GD.itemsize == 24 and np.dtype(F.GLYPH_DTYPE) == GD


def test_text_clear_and_single_glyph_known_values():
    m = np.array([[0, 128, 255]], np.uint8)
    g = np.zeros(1, GD)
    g[0] = (1, 0, 3, 1, 0, 0, (255, 255, 255, 255), MASK)
    # GpuOptimized, transparent background: coverage 0 leaves the clear, 255 gives opaque white; 128/255 of white over
    # transparent black: linear 0.50196 -> sRGB 188, alpha 128
    out = orc.render_text(5, 1, (0, 0, 0, 0), g, m, None, 0, orc.MODE_GPU_OPTIMIZED)
    assert out[0].tolist() == [[0, 0, 0, 0], [0, 0, 0, 0], [188, 188, 188, 128], [255, 255, 255, 255], [0, 0, 0, 0]]
    # the clear stores the premultiplied linear colour through the sRGB view: 128/255 x linear(255) = 0.50196 -> 188
    out = orc.render_text(2, 1, (255, 0, 0, 128), np.zeros(0, GD), None, None, 0, orc.MODE_GPU_OPTIMIZED)
    assert out[0].tolist() == [[188, 0, 0, 128]] * 2
    # CpuOptimized stores plain bytes: 128/255 x 255/255 -> 128
    out = orc.render_text(2, 1, (255, 0, 0, 128), np.zeros(0, GD), None, None, 0, orc.MODE_CPU_OPTIMIZED)
    assert out[0].tolist() == [[128, 0, 0, 128]] * 2
    # zero-sized text texture: one transparent pixel (text_renderer.rs:77-85)
    assert orc.render_text(0, 7, (9, 9, 9, 255), np.zeros(0, GD)).tolist() == [[[0, 0, 0, 0]]]


def test_text_color_mode_accurate_vs_web():
    """ColorMode::Accurate linearises the glyph colour before the blend, Web leaves the bytes: an opaque mid-grey glyph on
    an sRGB node texture stores 128 in Accurate mode (decode then encode) and encode(128/255) = 188 in Web mode"""
    m = np.full((1, 1), 255, np.uint8)
    g = np.zeros(1, GD)
    g[0] = (0, 0, 1, 1, 0, 0, (128, 128, 128, 255), MASK)
    assert orc.render_text(1, 1, (0, 0, 0, 0), g, m, None, 0, orc.MODE_GPU_OPTIMIZED)[0, 0].tolist() == [128, 128, 128, 255]
    assert orc.render_text(1, 1, (0, 0, 0, 0), g, m, None, 1, orc.MODE_GPU_OPTIMIZED)[0, 0].tolist() == [188, 188, 188, 255]


@pytest.mark.parametrize("mode", [orc.MODE_GPU_OPTIMIZED, orc.MODE_CPU_OPTIMIZED])
@pytest.mark.parametrize("color_mode", [0, 1])
def test_text_oracle_matches_independent_restatement(mode, color_mode):
    w, h = 40, 18
    m, c = soft_mask_atlas(64, 40, 1), color_atlas(48, 40, 2)
    g = random_glyphs(30, w, h, 48, 40, 5 + mode)
    bg = (40, 90, 200, 170)
    got = orc.render_text(w, h, bg, g, m, c, color_mode, mode)
    exp = text_ref(w, h, bg, g, m, c, color_mode, mode)
    assert np.array_equal(got, exp), f"{np.count_nonzero(got != exp)} bytes differ"


def test_text_painters_order_matters():
    m = np.full((8, 8), 200, np.uint8)
    g = np.zeros(2, GD)
    g[0] = (0, 0, 8, 8, 0, 0, (255, 0, 0, 255), MASK)
    g[1] = (0, 0, 8, 8, 0, 0, (0, 0, 255, 255), MASK)
    a = orc.render_text(8, 8, (0, 0, 0, 255), g, m, None)
    b = orc.render_text(8, 8, (0, 0, 0, 255), g[::-1].copy(), m, None)
    assert not np.array_equal(a, b) and a[0, 0, 2] > a[0, 0, 0] and b[0, 0, 0] > b[0, 0, 2]


# ---- GPU: product against the oracle -------------------------------------------------------------------------------
def _modes():
    import smelter_b200 as s
    return {orc.MODE_GPU_OPTIMIZED: s.RenderingMode.GpuOptimized, orc.MODE_CPU_OPTIMIZED: s.RenderingMode.CpuOptimized}


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [orc.MODE_GPU_OPTIMIZED, orc.MODE_CPU_OPTIMIZED])
@pytest.mark.parametrize("color_mode", [0, 1])
@pytest.mark.parametrize("geom", [(300, 70, 700), (33, 9, 40), (1, 1, 3), (1000, 120, 2000), (64, 64, 0)])
def test_text_product_matches_oracle(mode, color_mode, geom):
    import smelter_b200 as s
    w, h, n = geom
    m, c = soft_mask_atlas(256, 128, 3), color_atlas(128, 96, 4)
    g = random_glyphs(n, w, h, 128, 96, 17 + n)
    bg = s.RGBAColor(12, 200, 90, 140)
    r = s.Renderer(s.RendererOptions(rendering_mode=_modes()[mode]))
    got = r.render_text(w, h, bg, g, m, c, color_mode)
    exp = orc.render_text(w, h, (bg.r, bg.g, bg.b, bg.a), g, m, c, color_mode, mode)
    assert np.array_equal(got, exp), f"{np.count_nonzero(got != exp)} bytes differ"
    assert r.stats()["kernel_launches"] >= 1


@pytest.mark.gpu
def test_text_mask_only_and_argument_checks():
    import smelter_b200 as s
    r = s.Renderer()
    m = soft_mask_atlas(64, 32, 9)
    g = random_glyphs(50, 120, 40, 64, 32, 1, contents=(MASK,))
    got = r.render_text(120, 40, s.RGBAColor(0, 0, 0, 0), g, m, None)
    assert np.array_equal(got, orc.render_text(120, 40, (0, 0, 0, 0), g, m, None))
    g2 = random_glyphs(5, 120, 40, 64, 32, 1, contents=(COLOR,))
    with pytest.raises(s.RendererError):
        r.render_text(120, 40, s.RGBAColor(0, 0, 0, 0), g2, m, None)        # colour glyphs without a colour atlas
    g["content"][3] = 7
    with pytest.raises(s.RendererError):
        r.render_text(120, 40, s.RGBAColor(0, 0, 0, 0), g, m, None)
    assert r.render_text(0, 0, s.RGBAColor(1, 2, 3, 4), g[:0], None, None).tolist() == [[[0, 0, 0, 0]]]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [orc.MODE_GPU_OPTIMIZED, orc.MODE_CPU_OPTIMIZED])
def test_text_node_as_a_layer_of_a_scene(mode):
    """a Text component's node texture (rendered once per scene update) composited like any other child: 1:1 over video
    and scaled by a Rescaler -- the reference's `video_call_with_labels` shape (render_tests/view.rs)"""
    import smelter_b200 as s
    from tests import harness
    from tests.parity import assert_identical, run_case, yuv_frame
    w, h = 220, 48
    m = soft_mask_atlas(256, 64, 21)
    g = random_glyphs(60, w, h, 232, 44, 8, contents=(MASK,))
    g["color"][:] = (250, 250, 250, 255)
    r = s.Renderer(s.RendererOptions(rendering_mode=_modes()[mode]))
    label = r.render_text(w, h, s.RGBAColor(0, 0, 0, 120), g, m, None)
    assert np.array_equal(label, orc.render_text(w, h, (0, 0, 0, 120), g, m, None, 0, mode))
    V = s.ViewComponent
    frames = {"input_1": yuv_frame(harness.test_input(1, 640, 360), 640, 360),
              "text_1": s.Frame(s.FrameData.Rgba8(label), s.Resolution(w, h), 0.0)}
    text = lambda: s.InputStreamComponent(input_id="text_1")
    scene = V(background_color=s.RGBAColor(0x33, 0x33, 0x33, 255), children=[
        s.RescalerComponent(child=s.InputStreamComponent(input_id="input_1")),
        V(position=s.Position.Absolute(width=float(w), height=float(h), left=30.0, bottom=20.0), children=[text()]),
        s.RescalerComponent(position=s.Position.Absolute(width=330.0, height=72.0, right=10.0, top=10.0), child=text())])
    got, exp, _ = run_case(scene, frames, mode=_modes()[mode])
    assert_identical(got, exp, "text layer")
