"""Python restatement of the reference's render-test harness inputs
(integration-tests/src/render_tests/harness/input.rs:11-151).  Test infrastructure."""
import numpy as np

F = np.float32

COLOR_VARIANTS = [
    (255, 0, 0), (0, 255, 0), (255, 255, 0), (255, 0, 255), (0, 0, 255), (0, 255, 255),
    (255, 165, 0), (255, 255, 255), (128, 128, 128), (255, 128, 128), (128, 128, 255),
    (128, 255, 128), (255, 192, 203), (128, 0, 128), (165, 42, 42), (154, 205, 50),
    (255, 255, 224),
]


def rgb_to_yuv_f32(rgb):
    """RGBColor::to_yuv, smelter-render/src/scene/types.rs:28-42 (f32 arithmetic)."""
    r, g, b = (F(c) / F(255.0) for c in rgb)
    y = r * F(0.2126) + g * F(0.7152) + b * F(0.0722)
    u = r * F(-0.1146) + g * F(-0.3854) + b * F(0.5)
    v = r * F(0.5) + g * F(-0.4542) + b * F(-0.0458)
    k = F(16.0) / F(255.0)
    cl = lambda x: min(max(x, F(0.0)), F(1.0))
    return (cl(y * F(0.85882354) + k), cl((u + F(0.5)) * F(0.8784314) + k),
            cl((v + F(0.5)) * F(0.8784314) + k))


def _as_u8(x):
    """Rust `as u8`: truncate toward zero, saturate."""
    return np.uint8(min(max(int(np.trunc(x)), 0), 255))


def test_input(index, width=640, height=360):
    """TestInput::new_with_resolution (input.rs:58-110) -> (y, u, v) planes."""
    yc, uc, vc = rgb_to_yuv_f32(COLOR_VARIANTS[index])
    xs = np.arange(width)[None, :]
    ys = np.arange(height)[:, None]
    B, G = 18, 72
    border_x = (xs <= B) | ((xs <= width) & (xs >= width - B))
    border_y = (ys <= B) | ((ys <= height) & (ys >= height - B))
    grid = ((xs // G + ys // G) % 2) == 0
    dark = border_x | border_y | grid
    y_hi = _as_u8(min(max(yc, F(0)), F(1)) * F(255.0))
    y_lo = _as_u8(min(max(yc - F(0.2), F(0)), F(1)) * F(255.0))
    y = np.where(dark, y_lo, y_hi).astype(np.uint8)
    u_val = _as_u8((uc + uc + uc + uc) * F(64.0))
    v_val = _as_u8((vc + vc + vc + vc) * F(64.0))
    u = np.full((height // 2, width // 2), u_val, np.uint8)
    v = np.full((height // 2, width // 2), v_val, np.uint8)
    return y, u, v


def multiscale_grid(width, height):
    """TestInput::new_multiscale_grid (input.rs:116-151)."""
    y = np.full((height, width), 200, np.uint8)
    periods = np.array([21, 15, 12, 9, 7, 5, 4, 3])
    band_w = width // len(periods)
    xs = np.arange(width)
    band = np.minimum(xs // band_w, len(periods) - 1)
    per = periods[band]
    vline = (xs % per) < 2
    ys = np.arange(height)[:, None]
    hline = (ys % per[None, :]) < 2
    y[hline | vline[None, :]] = 30
    u = np.full((height // 2, width // 2), 128, np.uint8)
    v = np.full((height // 2, width // 2), 128, np.uint8)
    return y, u, v


def random_yuv420(seed, width, height):
    rng = np.random.default_rng(seed)
    y = rng.integers(16, 236, (height, width), dtype=np.uint8)
    u = rng.integers(16, 241, (height // 2, width // 2), dtype=np.uint8)
    v = rng.integers(16, 241, (height // 2, width // 2), dtype=np.uint8)
    return y, u, v


def smooth_yuv420(seed, width, height):
    """Band-limited random content (more video-like than white noise)."""
    rng = np.random.default_rng(seed)
    def plane(w, h, lo, hi):
        g = rng.random((h // 16 + 2, w // 16 + 2))
        g = np.kron(g, np.ones((16, 16)))[:h, :w]
        g = g + 0.08 * rng.random((h, w))
        g = (g - g.min()) / (g.max() - g.min() + 1e-9)
        return (lo + g * (hi - lo)).astype(np.uint8)
    return plane(width, height, 16, 235), plane(width // 2, height // 2, 16, 240), plane(width // 2, height // 2, 16, 240)
