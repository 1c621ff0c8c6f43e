"""Names the re-typed scene catalogue (tests/golden/ref_scenes.py) refers to: the reference's scene types and its
render-test harness (integration-tests/src/render_tests/harness/{test_case,input}.rs), mapped onto the Python mirror of
the C ABI.  `TestRunner` only RECORDS a test: inputs, resolution, rendering mode and the sequence of scene updates and
snapshots; tests replay the recording against the product / the oracle / the independent layout engine."""
import smelter_b200 as s

RGBAColor = s.RGBAColor
BoxShadow = s.BoxShadow
BorderRadius = s.BorderRadius
HorizontalAlign = s.HorizontalAlign
VerticalAlign = s.VerticalAlign
Overflow = s.Overflow
ViewChildrenDirection = s.ViewChildrenDirection
RescaleMode = s.RescaleMode
RenderingMode = s.RenderingMode


DEFAULT_RESOLUTION = s.Resolution(640, 360)   # render_tests/harness.rs


class Unsupported(Exception):
    """the test uses a component outside the compositor hot path (text / image / shader)"""


def rmap(x, fn):
    """Option::map or Iterator::map"""
    if x is None:
        return None
    if isinstance(x, (range, list, tuple)) or hasattr(x, "__iter__") and not isinstance(x, str):
        return [fn(v) for v in x]
    return fn(x)


def Padding(top=0.0, right=0.0, bottom=0.0, left=0.0):
    return s.Padding(top, right, bottom, left)


def Resolution(width, height):
    return s.Resolution(int(width), int(height))   # usize arithmetic in the Rust tests (e.g. height / 2) truncates


def InputId(x):
    return x


def ComponentId(x):
    return x


def RendererId(x):
    return x


class Duration:
    ZERO = 0.0

    @staticmethod
    def from_millis(ms):
        return ms / 1000.0

    @staticmethod
    def from_secs(sec):
        return float(sec)

    @staticmethod
    def from_secs_f64(sec):
        return float(sec)


class HorizontalPosition:
    LeftOffset = staticmethod(lambda v: ("left", v))
    RightOffset = staticmethod(lambda v: ("right", v))


class VerticalPosition:
    TopOffset = staticmethod(lambda v: ("top", v))
    BottomOffset = staticmethod(lambda v: ("bottom", v))


def AbsolutePosition(width=None, height=None, position_horizontal=("left", 0.0), position_vertical=("top", 0.0),
                     rotation_degrees=0.0):
    kw = {position_horizontal[0]: position_horizontal[1], position_vertical[0]: position_vertical[1]}
    return s.Position.Absolute(width=width, height=height, rotation_degrees=rotation_degrees, **kw)


class Position:
    Static = staticmethod(lambda width=None, height=None: s.Position.Static(width=width, height=height))
    Absolute = staticmethod(lambda p: p)


class InterpolationKind:
    Linear = s.InterpolationKind.Linear
    Bounce = s.InterpolationKind.Bounce
    CubicBezier = staticmethod(lambda x1, y1, x2, y2: s.InterpolationKind.CubicBezier(x1, y1, x2, y2))


def Transition(duration=0.0, interpolation_kind=s.InterpolationKind.Linear, should_interrupt=False):
    return s.Transition(duration=duration, interpolation_kind=interpolation_kind, should_interrupt=should_interrupt)


def InputStreamComponent(id=None, input_id=""):
    return s.InputStreamComponent(input_id=input_id, id=id)


class _View:
    def __call__(self, **kw):
        return s.ViewComponent(**kw)

    @staticmethod
    def default():
        return s.ViewComponent()


ViewComponent = _View()


def RescalerComponent(**kw):
    return s.RescalerComponent(**kw)


def TilesComponent(**kw):
    return s.TilesComponent(**kw)


def _unsupported(*a, **kw):
    raise Unsupported()


TextComponent = ImageComponent = ShaderComponent = _unsupported


class TextDimensions:
    Fitted = Fixed = FittedColumn = staticmethod(lambda **kw: None)


class Component:
    View = Rescaler = Tiles = InputStream = staticmethod(lambda c: c)
    Text = Image = Shader = WebView = staticmethod(_unsupported)


# ---- harness ----------------------------------------------------------------------------------------------------------
class TestInput:
    """harness/input.rs: (index, resolution, pattern)"""
    __test__ = False

    def __init__(self, index, resolution=None, pattern="checker"):
        self.index, self.resolution, self.pattern = index, resolution or s.Resolution(640, 360), pattern
        self.name = f"input_{index}"

    @staticmethod
    def new(index):
        return TestInput(index)

    @staticmethod
    def new_with_resolution(index, resolution):
        return TestInput(index, resolution)

    @staticmethod
    def new_multiscale_grid(index, resolution):
        return TestInput(index, resolution, "multiscale_grid")


class TestRunner:
    __test__ = False

    def __init__(self, module, name):
        self.module, self.name = module, name
        self.inputs = []
        self.resolution = s.Resolution(640, 360)          # harness/test_case.rs DEFAULT_RESOLUTION
        self.mode = s.RenderingMode.GpuOptimized
        self.steps = []                                    # ("update", scene) | ("snapshot", pts seconds)

    @staticmethod
    def new(module, name):
        return TestRunner(module, name)

    def with_inputs(self, inputs):
        self.inputs = list(inputs)
        return self

    def with_resolution(self, resolution):
        self.resolution = resolution
        return self

    def with_rendering_mode(self, mode):
        self.mode = mode
        return self

    def with_renderers(self, renderers):
        raise Unsupported()

    def update_scene(self, scene):
        self.steps.append(("update", scene))

    def snapshot(self, pts):
        self.steps.append(("snapshot", float(pts)))

    def render(self, pts):
        self.steps.append(("render", float(pts)))

    def finish(self):
        return self


def record(test_fn):
    """run a re-typed test function; returns its TestRunner recording, or None when it is out of scope"""
    try:
        return test_fn()
    except Unsupported:
        return None
