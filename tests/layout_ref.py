"""Independent restatement of the reference's layout engine, in Python / numpy float32.  Test infrastructure.

Typed from the Rust sources (NOT from smelter_b200/csrc/scene.cpp), so that the product's host engine is checked
against something other than itself:

  scene/layout.rs:95-236                      update_state, layout_content, layout_absolute_position_child
  scene/view_component/layout.rs:31-285       View: static / absolute children, overflow, padding, borders
  scene/rescaler_component/layout.rs:14-161   Rescaler: fit / fill, alignment
  scene/tiles_component/tiles.rs:29-165       Tiles: rows x columns search, tile size, positions
  scene/tiles_component/layout.rs:10-128      layout_tiles, fit_into_tile
  scene/types.rs:109-160                      BorderRadius clip / + / - / * / '/'
  transformations/layout/flatten.rs:10-390    flatten, should_render, fix_final_render_layout, masks
  scene/transition.rs:39-106, scene/**/interpolation.rs   transitions (View / Rescaler / Tiles), see StatefulScene

`layouts(scene, resolution, input_resolutions)` gives the flattened RenderLayout list of a freshly registered scene
(no previous state); `StatefulScene` carries state across update_scene calls the way scene_state.rs does, for the
transition tests.  Every quantity is an np.float32 and every operation is written in the reference's order.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

import smelter_b200 as s

F = np.float32
ZERO, ONE, TWO = F(0.0), F(1.0), F(2.0)


def f(x):
    return F(x)


def fmax(a, b):   # f32::max: NaN-ignoring
    a, b = F(a), F(b)
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a > b else b


def fmin(a, b):
    a, b = F(a), F(b)
    if np.isnan(a):
        return b
    if np.isnan(b):
        return a
    return a if a < b else b


# ---- scene/types.rs ----------------------------------------------------------------------------------------------
@dataclass
class Radius:
    tl: np.float32 = ZERO
    tr: np.float32 = ZERO
    br: np.float32 = ZERO
    bl: np.float32 = ZERO

    @staticmethod
    def of(b):
        return Radius(F(b.top_left), F(b.top_right), F(b.bottom_right), F(b.bottom_left))

    def clip_to_size(self, w, h):   # types.rs:109-117
        m = fmax(ZERO, fmin(w, h) / TWO)
        c = lambda v: fmin(fmax(v, ZERO), m)    # f32::clamp
        return Radius(c(self.tl), c(self.tr), c(self.br), c(self.bl))

    def mul(self, k):
        k = F(k)
        return Radius(self.tl * k, self.tr * k, self.br * k, self.bl * k)

    def div(self, k):               # self * (1.0 / rhs)
        return self.mul(ONE / F(k))

    def add(self, k):               # max(x + rhs, 0)
        k = F(k)
        return Radius(fmax(self.tl + k, ZERO), fmax(self.tr + k, ZERO), fmax(self.br + k, ZERO), fmax(self.bl + k, ZERO))

    def sub(self, k):
        return self.add(-F(k))

    def tup(self):
        return (self.tl, self.tr, self.br, self.bl)


@dataclass
class Mask:
    radius: Radius
    top: np.float32
    left: np.float32
    width: np.float32
    height: np.float32


@dataclass
class Crop:
    top: np.float32
    left: np.float32
    width: np.float32
    height: np.float32


@dataclass
class Nested:                        # transformations/layout.rs NestedLayout
    top: np.float32
    left: np.float32
    width: np.float32
    height: np.float32
    rotation: np.float32 = ZERO
    scale_x: np.float32 = ONE
    scale_y: np.float32 = ONE
    crop: Optional[Crop] = None
    mask: Optional[Mask] = None
    content: tuple = ("none",)      # ("none",) | ("color", rgba) | ("child", index, w, h)
    border_width: np.float32 = ZERO
    border_color: tuple = (0, 0, 0, 0)
    border_radius: Radius = field(default_factory=Radius)
    box_shadow: list = field(default_factory=list)
    children: list = field(default_factory=list)
    child_nodes_count: int = 0


@dataclass
class Render:                        # RenderLayout
    top: np.float32
    left: np.float32
    width: np.float32
    height: np.float32
    rotation: np.float32
    border_radius: Radius
    masks: List[Mask]
    kind: str                        # "color" | "child" | "shadow"
    color: tuple = (0, 0, 0, 0)
    border_color: tuple = (0, 0, 0, 0)
    border_width: np.float32 = ZERO
    blur_radius: np.float32 = ZERO
    index: int = -1
    crop: Optional[Crop] = None


def rgba(c):
    return (int(c.r), int(c.g), int(c.b), int(c.a))


# ---- stateful component tree (scene_state.rs, *_component.rs) ----------------------------------------------------------
NS = 1_000_000_000


def secs_f64(ns):
    """Duration::as_secs_f64"""
    return float(ns // NS) + float(ns % NS) / 1e9


def to_ns(seconds):
    return int(round(float(seconds) * 1e9))


def lerp32(a, b, st):
    """ContinuousValue for f32: (start as f64 + (end - start) * state) as f32"""
    return F(float(a) + ((float(b) - float(a)) * st))


def lerp_opt(a, b, st):
    if a is not None and b is not None:
        return lerp32(a, b, st)
    return b


def bounce_easing(t):
    n1, d1 = 7.5625, 2.75
    if t < (1.0 / d1):
        return n1 * t * t
    if t < (2.0 / d1):
        return n1 * (t - 1.5 / d1) * (t - 1.5 / d1) + 0.75
    if t < (2.5 / d1):
        return n1 * (t - 2.25 / d1) * (t - 2.25 / d1) + 0.9375
    return n1 * (t - 2.625 / d1) * (t - 2.625 / d1) + 0.984375


def cubic_bezier_easing(progress, x1, y1, x2, y2):
    """transition/cubic_bezier.rs"""
    import math
    EPS = 1e-7
    close = lambda a, b: abs(a - b) < EPS

    def clamp_root(v):
        if v < 0.0:
            return 0.0 if v >= -EPS else math.nan
        if v > 1.0:
            return 1.0 if v <= 1.0 + EPS else math.nan
        return v

    def cbrt(v):
        return math.copysign(abs(v) ** (1.0 / 3.0), v) if v != 0 else 0.0

    def first_root(p0, p1, p2, p3):
        a = 3.0 * (p0 - 2.0 * p1 + p2)
        b = 3.0 * (p1 - p0)
        c = p0
        d = -p0 + 3.0 * (p1 - p2) + p3
        if close(d, 0.0):
            if close(a, 0.0):
                if close(b, 0.0):
                    return math.nan
                return clamp_root(-c / b)
            disc = b * b - 4.0 * a * c
            q = math.sqrt(disc) if disc >= 0 else math.nan
            a2 = 2.0 * a
            root = clamp_root((q - b) / a2) if not math.isnan(q) else math.nan
            if not math.isnan(root):
                return root
            return clamp_root((-b - q) / a2) if not math.isnan(q) else math.nan
        a, b, c = a / d, b / d, c / d
        o3 = (3.0 * b - a ** 2) / 9.0
        q2 = (2.0 * a ** 3 - 9.0 * a * b + 27.0 * c) / 54.0
        a3 = a / 3.0
        disc = q2 ** 2 + o3 ** 3
        if disc < 0.0:
            mp33 = -(o3 ** 3)
            r = math.sqrt(mp33)
            cos_phi = min(max(-q2 / r, -1.0), 1.0)
            phi = math.acos(cos_phi)
            t1 = 2.0 * cbrt(r)
            for k in (0.0, 2.0 * math.pi, 4.0 * math.pi):
                root = clamp_root(t1 * math.cos((phi + k) / 3.0) - a3)
                if not math.isnan(root) or k == 4.0 * math.pi:
                    return root
        if disc == 0.0:
            u1 = -cbrt(q2)
            root = clamp_root(2.0 * u1 - a3)
            if not math.isnan(root):
                return root
            return clamp_root(-u1 - a3)
        sd = math.sqrt(disc)
        return clamp_root(cbrt(-q2 + sd) - cbrt(q2 + sd) - a3)

    if close(progress, 0.0):
        return 0.0
    if close(progress, 1.0):
        return 1.0
    t = first_root(-progress, x1 - progress, x2 - progress, 1.0 - progress)
    if math.isnan(t):
        return 1.0
    a = 1.0 / 3.0 + (y1 - y2)
    b = y2 - 2.0 * y1
    c = y1
    return min(max(3.0 * ((a * t + b) * t + c) * t, 0.0), 1.0)


def kind_state(kind, t):
    if kind.kind == 0:
        return t
    if kind.kind == 1:
        return bounce_easing(t)
    return cubic_bezier_easing(t, kind.x1, kind.y1, kind.x2, kind.y2)


class TransitionState:                 # scene/transition.rs
    def __init__(self, offset, start_ns, duration_ns, kind):
        self.offset, self.start_ns, self.duration_ns, self.kind = offset, start_ns, duration_ns, kind

    @staticmethod
    def new(current, previous, props_changed, interrupt, last_ns):
        """current: smelter Transition or None; previous: TransitionState or None"""
        def from_options(t):
            return TransitionState((0.0, 0.0), last_ns, to_ns(t.duration), t.interpolation_kind)
        if previous is not None and not previous.is_finished(last_ns):
            if props_changed and interrupt:
                return from_options(current) if current is not None else None
            remaining = max(previous.start_ns + previous.duration_ns - last_ns, 0)
            progress_offset = 1.0 - (secs_f64(remaining) / secs_f64(previous.duration_ns))
            state_offset = kind_state(previous.kind, progress_offset)
            return TransitionState((progress_offset, state_offset), last_ns, remaining,
                                   current.interpolation_kind if current is not None else previous.kind)
        if props_changed:
            return from_options(current) if current is not None else None
        return None

    def state(self, pts_ns):
        with np.errstate(all="ignore"):
            d = secs_f64(self.duration_ns)
            num = secs_f64(pts_ns) - secs_f64(self.start_ns)
            progress = num / d if d != 0.0 else (float("nan") if num == 0.0 else float("inf") if num > 0 else float("-inf"))
        progress = self.offset[0] + progress * (1.0 - self.offset[0])
        progress = min(max(progress, 0.0), 1.0) if progress == progress else progress   # f64::clamp keeps NaN
        st = kind_state(self.kind, progress)
        den = 1.0 - self.offset[1]
        return (st - self.offset[1]) / den if den != 0.0 else float("nan")

    def is_finished(self, now_ns):
        return self.start_ns + self.duration_ns <= now_ns


def pos_of(p):
    if p.absolute:
        hor = ("right", F(p.right)) if p.right is not None else ("left", F(p.left or 0.0))
        ver = ("bottom", F(p.bottom)) if p.bottom is not None else ("top", F(p.top or 0.0))
        return ("absolute", None if p.width is None else F(p.width), None if p.height is None else F(p.height), hor, ver,
                F(p.rotation_degrees))
    return ("static", None if p.width is None else F(p.width), None if p.height is None else F(p.height))


def pos_lerp(a, b, st):               # components/interpolation.rs
    if a[0] == "static" and b[0] == "static":
        return ("static", lerp_opt(a[1], b[1], st), lerp_opt(a[2], b[2], st))
    if a[0] == "absolute" and b[0] == "absolute":
        hor = (b[3][0], lerp32(a[3][1], b[3][1], st)) if a[3][0] == b[3][0] else b[3]
        ver = (b[4][0], lerp32(a[4][1], b[4][1], st)) if a[4][0] == b[4][0] else b[4]
        return ("absolute", lerp_opt(a[1], b[1], st), lerp_opt(a[2], b[2], st), hor, ver, lerp32(a[5], b[5], st))
    return b


def pos_grow(p, dx, dy):              # with_border / with_padding: width + dx, height + dy when present
    w = None if p[1] is None else p[1] + dx
    h = None if p[2] is None else p[2] + dy
    return (p[0], w, h) + tuple(p[3:])


@dataclass
class Shadow:
    offset_x: np.float32
    offset_y: np.float32
    blur_radius: np.float32
    color: object


def params_of(c):
    """the interpolated part of ViewComponentParam / RescalerComponentParam"""
    pd = getattr(c, "padding", None)
    return {"position": pos_of(c.position), "border_width": F(c.border_width), "border_radius": Radius.of(c.border_radius),
            "box_shadow": [Shadow(F(b.offset_x), F(b.offset_y), F(b.blur_radius), b.color) for b in c.box_shadow],
            "padding": None if pd is None else (F(pd.top), F(pd.right), F(pd.bottom), F(pd.left))}


def params_lerp(a, b, st):
    r0, r1 = a["border_radius"], b["border_radius"]
    sh = [Shadow(lerp32(x.offset_x, y.offset_x, st), lerp32(x.offset_y, y.offset_y, st), lerp32(x.blur_radius, y.blur_radius, st), y.color)
          for x, y in zip(a["box_shadow"], b["box_shadow"])] + b["box_shadow"][min(len(a["box_shadow"]), len(b["box_shadow"])):]
    pad = None if b["padding"] is None else tuple(lerp32(x, y, st) for x, y in zip(a["padding"], b["padding"]))
    return {"position": pos_lerp(a["position"], b["position"], st), "border_width": lerp32(a["border_width"], b["border_width"], st),
            "border_radius": Radius(lerp32(r0.tl, r1.tl, st), lerp32(r0.tr, r1.tr, st), lerp32(r0.br, r1.br, st), lerp32(r0.bl, r1.bl, st)),
            "box_shadow": sh, "padding": pad}


def comparable(c):
    """PartialEq of the *ComponentParam structs: every field but children / child / transition"""
    d = {k: v for k, v in vars(c).items() if k not in ("children", "child", "transition")}
    return (type(c).__name__, repr(sorted(d.items(), key=lambda kv: kv[0])))


class SNode:
    """StatefulComponent"""

    def __init__(self, comp, ctx):
        self.comp = comp
        self.kind = ("input" if isinstance(comp, s.InputStreamComponent) else "view" if isinstance(comp, s.ViewComponent)
                     else "rescaler" if isinstance(comp, s.RescalerComponent) else "tiles")
        prev = ctx["prev"].get(comp.id) if getattr(comp, "id", None) is not None else None
        if prev is not None and prev.kind != self.kind:
            prev = None
        last = ctx["last_ns"]
        if self.kind == "input":
            r = ctx["resolutions"].get(comp.input_id)
            self.size = (F(r[0]), F(r[1])) if r is not None else (ZERO, ZERO)
            self.children = []
            return
        kids = [comp.child if comp.child is not None else s.ViewComponent()] if self.kind == "rescaler" else list(comp.children)
        if self.kind in ("view", "rescaler"):
            self.start = prev.params(last) if prev is not None else None
            self.end = params_of(comp)
            changed = prev is not None and comparable(prev.comp) != comparable(comp)
        else:
            self.start = prev.last_layout if prev is not None else None
            self.last_layout = prev.last_layout if prev is not None else None
            changed = False
            if prev is not None:
                ids_a = [getattr(k.comp, "id", None) for k in prev.children]
                ids_b = [getattr(k, "id", None) for k in kids]
                changed = comparable(prev.comp) != comparable(comp) or ids_a != ids_b
        t = comp.transition
        self.transition = TransitionState.new(t, prev.transition if prev is not None else None, changed,
                                              bool(t.should_interrupt) if t is not None else False, last)
        self.children = [SNode(k, ctx) for k in kids]

    def params(self, pts_ns):          # view() / transition_snapshot()
        if self.transition is None or self.start is None:
            return self.end
        return params_lerp(self.start, self.end, self.transition.state(pts_ns))

    def with_id(self, out):
        if getattr(self.comp, "id", None) is not None:
            out[self.comp.id] = self
        for k in self.children:
            k.with_id(out)
        return out

    def clone(self):
        import copy
        c = copy.copy(self)
        c.children = [k.clone() for k in self.children]
        return c

    def node_children(self):           # layout.rs:84-93
        out = []
        for k in self.children:
            out += [k] if k.kind == "input" else k.node_children()
        return out


class Engine:
    """NestedLayout of a stateful tree at one pts"""

    def __init__(self, pts_ns):
        self.pts = pts_ns

    def is_layout(self, n):
        return n.kind != "input"

    # -- position / sizes ----------------------------------------------------------------------------------------
    def position(self, n):           # external position: includes border (and padding for View)
        if n.kind == "tiles":
            c = n.comp
            return ("static", None if c.width is None else F(c.width), None if c.height is None else F(c.height))
        P = n.params(self.pts)
        bw2 = TWO * P["border_width"]
        p = pos_grow(P["position"], bw2, bw2)
        if n.kind == "view":
            t, r, b, l = P["padding"]
            p = pos_grow(p, l + r, t + b)
        return p

    def width(self, n):              # scene.rs:102-114
        return n.size[0] if n.kind == "input" else self.position(n)[1]

    def height(self, n):
        return n.size[1] if n.kind == "input" else self.position(n)[2]

    def layout_content(self, n, index):   # layout.rs:133-157
        if self.is_layout(n):
            return ("none",)
        return ("child", index, n.size[0], n.size[1])

    def update_state(self, n, sizes):      # layout.rs:95-131; sizes: per node child, (w, h) or None
        i = 0
        for k in n.children:
            if k.kind == "input":
                r = sizes[i]
                k.size = (F(r[0]), F(r[1])) if r is not None else (ZERO, ZERO)
                i += 1
            else:
                cnt = len(k.node_children())
                self.update_state(k, sizes[i:i + cnt])
                i += cnt

    # -- dispatch ------------------------------------------------------------------------------------------------
    def layout(self, n, w, h):
        if n.kind == "view":
            return self.view_layout(n, F(w), F(h))
        if n.kind == "rescaler":
            return self.rescaler_layout(n, F(w), F(h))
        return self.tiles_layout(n, F(w), F(h))

    # -- layout.rs:159-236 -----------------------------------------------------------------------------------------
    def absolute_child(self, child, pos, pw, ph):
        _, pwid, phei, hor, ver, rot = pos
        width = pwid if pwid is not None else pw
        height = phei if phei is not None else ph
        top = ver[1] if ver[0] == "top" else (ph - ver[1]) - height
        left = hor[1] if hor[0] == "left" else (pw - hor[1]) - width
        content = self.layout_content(child, 0)
        if self.is_layout(child):
            cl = self.layout(child, width, height)
            cnt = cl.child_nodes_count + (1 if content[0] == "child" else 0)
            return Nested(top, left, width, height, rotation=rot, content=content, child_nodes_count=cnt, children=[cl])
        return Nested(top, left, width, height, rotation=rot, content=content, child_nodes_count=1 if content[0] == "child" else 0)

    # -- view_component/layout.rs ---------------------------------------------------------------------------------
    def view_layout(self, n, sw, sh):
        c = n.comp
        P = n.params(self.pts)
        bw = P["border_width"]
        pt, pr, pb, pl = P["padding"]
        cw = fmax(sw - TWO * bw, ZERO)
        chh = fmax(sh - TWO * bw, ZERO)
        br = P["border_radius"].clip_to_size(sw, sh)
        kids = n.children
        row = c.direction == s.ViewChildrenDirection.Row
        static_child_size = self.static_child_size(row, (pt, pr, pb, pl), cw, chh, kids)
        ov = c.overflow
        if ov == s.Overflow.Visible:
            scale, mask = ONE, None
        else:
            scale = ONE if ov == s.Overflow.Hidden else self.scale_factor_for_overflow_fit(row, cw, chh, kids)
            mask = Mask(br.sub(bw), bw, bw, cw, chh)
        static_offset = bw / scale
        out = []
        for ch in kids:
            if self.is_layout(ch):
                pos = self.position(ch)
            else:
                pos = ("static", self.width(ch), self.height(ch))
            if pos[0] == "static":
                _, w_, h_ = pos
                pbw = bw / scale
                if row:
                    width = w_ if w_ is not None else static_child_size
                    height = h_ if h_ is not None else chh - (pt + pb)
                    top = pbw + pt
                    left = static_offset + pl
                    static_offset = static_offset + width
                else:
                    height = h_ if h_ is not None else static_child_size
                    width = w_ if w_ is not None else cw - (pl + pr)
                    top = static_offset + pt
                    left = pbw + pl
                    static_offset = static_offset + height
                if self.is_layout(ch):
                    cl = self.layout(ch, width, height)
                    out.append(Nested(top, left, width, height, content=("none",), child_nodes_count=cl.child_nodes_count, children=[cl]))
                else:
                    out.append(Nested(top, left, width, height, content=self.layout_content(ch, 0), child_nodes_count=1))
            else:
                out.append(self.absolute_child(ch, pos, sw, sh))
        return Nested(ZERO, ZERO, sw, sh, scale_x=scale, scale_y=scale, mask=mask, content=("color", rgba(c.background_color)),
                      child_nodes_count=sum(l.child_nodes_count for l in out), children=out, border_width=bw,
                      border_color=rgba(c.border_color), border_radius=br, box_shadow=list(P["box_shadow"]))

    def static_children(self, kids):
        return [k for k in kids if not self.is_layout(k) or self.position(k)[0] == "static"]

    def sum_static_children_sizes(self, row, kids):
        acc = ZERO
        for k in self.static_children(kids):
            v = self.width(k) if row else self.height(k)
            acc = acc + (v if v is not None else ZERO)
        return acc

    def static_child_size(self, row, pad, cw, chh, kids):
        pt, pr, pb, pl = pad
        max_size = cw - (pl + pr) if row else chh - (pt + pb)
        unknown = sum(1 for k in self.static_children(kids) if (self.width(k) if row else self.height(k)) is None)
        total = self.sum_static_children_sizes(row, kids)
        if unknown == 0:
            return ZERO
        return fmax(ZERO, (max_size - total) / F(unknown))

    def scale_factor_for_overflow_fit(self, row, cw, chh, kids):
        sum_size = fmax(self.sum_static_children_sizes(row, kids), F(0.000000001))
        max_size, max_alt = (cw, chh) if row else (chh, cw)
        best = None
        for k in self.static_children(kids):
            v = self.height(k) if row else self.width(k)
            v = v if v is not None else ZERO
            if best is None or not (best > v):     # Iterator::max_by keeps the LAST maximum
                best = v
        alt = fmax(best if best is not None else ZERO, F(0.000000001))
        return fmin(ONE, fmin(max_size / sum_size, max_alt / alt))

    # -- rescaler_component/layout.rs -----------------------------------------------------------------------------
    def rescaler_layout(self, n, sw, sh):
        c = n.comp
        P = n.params(self.pts)
        bw = P["border_width"]
        cw = fmax(sw - (TWO * bw), ZERO)
        chh = fmax(sh - (TWO * bw), ZERO)
        child = n.children[0]
        w_, h_ = self.width(child), self.height(child)
        br = P["border_radius"].clip_to_size(sw, sh)
        if w_ is None and h_ is None:
            scale = ONE
        elif w_ is None:
            scale = chh / h_
        elif h_ is None:
            scale = cw / w_
        elif c.mode == s.RescaleMode.Fit:
            scale = fmin(cw / w_, chh / h_)
        else:
            scale = fmax(cw / w_, chh / h_)
        if self.is_layout(child):
            cl = self.layout(child, w_ if w_ is not None else cw / scale, h_ if h_ is not None else chh / scale)
            content, kids, cnt = ("none",), [cl], cl.child_nodes_count
        else:
            content, kids, cnt = self.layout_content(child, 0), [], 1
        va, ha = c.vertical_align, c.horizontal_align
        if va == s.VerticalAlign.Top or h_ is None:
            top = ZERO
        elif va == s.VerticalAlign.Bottom:
            top = chh - (h_ * scale)
        else:
            top = (chh - (h_ * scale)) / TWO
        if ha == s.HorizontalAlign.Left or w_ is None:
            left = ZERO
        elif ha == s.HorizontalAlign.Right:
            left = cw - (w_ * scale)
        else:
            left = (cw - (w_ * scale)) / TWO
        width = w_ * scale if w_ is not None else cw
        height = h_ * scale if h_ is not None else chh
        inner = Nested(top + bw, left + bw, width, height, scale_x=scale, scale_y=scale, content=content,
                       child_nodes_count=cnt, children=kids)
        return Nested(ZERO, ZERO, cw + (bw * TWO), chh + (bw * TWO), mask=Mask(br.sub(bw), bw, bw, cw, chh), content=("none",),
                      children=[inner], child_nodes_count=cnt, border_width=bw, border_color=rgba(c.border_color),
                      border_radius=br, box_shadow=list(P["box_shadow"]))

    # -- tiles_component/tiles.rs + layout.rs ---------------------------------------------------------------------
    def tile_size(self, c, rows, cols, lw, lh):
        pad, mar = F(c.padding), F(c.margin)
        x_padding = F(cols) * TWO * pad
        y_padding = F(rows) * TWO * pad
        x_margin = (F(cols) + ONE) * mar
        y_margin = (F(rows) + ONE) * mar
        ax, ay = F(c.tile_aspect_ratio[0]), F(c.tile_aspect_ratio[1])
        x_scale = fmax(lw - x_padding - x_margin, ZERO) / F(cols) / ax
        y_scale = fmax(lh - y_padding - y_margin, ZERO) / F(rows) / ay
        scale = x_scale if x_scale < y_scale else y_scale
        return ax * scale, ay * scale

    def end_tiles(self, n, lw, lh):
        c = n.comp
        kids = n.children
        cnt = len(kids)
        if cnt == 0:
            return []
        best = (1, cnt)
        best_w = ZERO
        for rows in range(1, cnt + 1):
            cols = -(-cnt // rows)
            tw, _ = self.tile_size(c, rows, cols, lw, lh)
            if tw > best_w:
                best, best_w = (rows, cols), tw
        rows, cols = best
        tw, th = self.tile_size(c, rows, cols, lw, lh)
        pad, mar = F(c.padding), F(c.margin)
        add_y = lh - (th + TWO * pad) * F(rows) - (mar * (F(rows) + ONE))
        va = c.vertical_align
        if va == s.VerticalAlign.Top:
            add_top, just_y = ZERO, ZERO
        elif va == s.VerticalAlign.Center:
            add_top, just_y = add_y / TWO, ZERO
        elif va == s.VerticalAlign.Bottom:
            add_top, just_y = add_y, ZERO
        else:
            add_top, just_y = ZERO, add_y / (F(rows) + ONE)
        out = []
        top = add_top + just_y + pad + mar
        for row in range(rows):
            in_row = cols if row < rows - 1 else cnt - (rows - 1) * cols
            add_x = lw - (tw + TWO * pad) * F(in_row) - (mar * (F(in_row) + ONE))
            ha = c.horizontal_align
            if ha == s.HorizontalAlign.Left:
                add_left, just_x = ZERO, ZERO
            elif ha == s.HorizontalAlign.Right:
                add_left, just_x = add_x, ZERO
            elif ha == s.HorizontalAlign.Justified:
                add_left, just_x = ZERO, add_x / F(in_row + 1)
            else:
                add_left, just_x = add_x / TWO, ZERO
            left = add_left + just_x + mar + pad
            for _ in range(in_row):
                out.append([top, left, tw, th])
                left = left + (tw + mar + pad * TWO + just_x)
            top = top + (th + mar + pad * TWO + just_y)
        # ids: component id when present, else running index of the id-less children (tiles.rs:44-52)
        idx = 0
        tiles = []
        for t, k in zip(out, kids):
            cid = getattr(k.comp, "id", None)
            if cid is not None:
                tid = ("id", cid)
            else:
                tid = ("index", idx)
                idx += 1
            tiles.append({"id": tid, "top": t[0], "left": t[1], "width": t[2], "height": t[3]})
        return tiles

    def tiles(self, n, sw, sh):       # StatefulTilesComponent::tiles
        end = self.end_tiles(n, sw, sh)
        if n.start is None or n.transition is None:
            return end
        start, (stw, sth) = n.start
        k = fmin(sw / stw, sh / sth)           # resize_tiles
        start = [None if t is None else {"id": t["id"], "top": t["top"] * k, "left": t["left"] * k, "width": t["width"] * k,
                                         "height": t["height"] * k} for t in start]
        st = n.transition.state(self.pts)
        # tiles_component/interpolation.rs
        start_ids = {}
        for i, t in enumerate(start):
            if t is not None:
                start_ids[t["id"]] = i
        end_ids = {t["id"] for t in end if t is not None}
        if st >= 1.0:
            return end
        out = []
        for t in end:
            if t is None:
                out.append(None)
                continue
            old = start[start_ids[t["id"]]] if t["id"] in start_ids else None
            if old is not None:
                out.append({"id": t["id"], "top": lerp32(old["top"], t["top"], st), "left": lerp32(old["left"], t["left"], st),
                            "width": lerp32(old["width"], t["width"], st), "height": lerp32(old["height"], t["height"], st)})
                continue
            tol = F(0.001)
            same = next((x for x in start if x is not None and abs(x["top"] - t["top"]) <= tol and abs(x["left"] - t["left"]) <= tol and
                         abs(x["width"] - t["width"]) <= tol and abs(x["height"] - t["height"]) <= tol), None)
            if same is not None:
                out.append(None if same["id"] in end_ids else dict(t))
            else:
                out.append(None)
        return out

    def tiles_layout(self, n, sw, sh):
        c = n.comp
        tiles = self.tiles(n, sw, sh)
        out = []
        for ch, tile in zip(n.children, tiles):
            if tile is None:     # child_nodes_placeholder
                cnt = len(ch.node_children()) if self.is_layout(ch) else 1
                out.append(Nested(ZERO, ZERO, ZERO, ZERO, content=("none",), child_nodes_count=cnt))
                continue
            if self.is_layout(ch):
                cl = self.layout(ch, tile["width"], tile["height"])
                out.append(Nested(tile["top"], tile["left"], tile["width"], tile["height"], content=("none",),
                                  child_nodes_count=cl.child_nodes_count, children=[cl]))
            else:
                w_, h_ = self.width(ch), self.height(ch)
                top, left, tw, th = tile["top"], tile["left"], tile["width"], tile["height"]
                if w_ is not None and h_ is not None:   # fit_into_tile
                    with np.errstate(all="ignore"):
                        sfac = fmin(tw / w_, th / h_)
                        top_off = (th - sfac * h_) / TWO
                        left_off = (tw - sfac * w_) / TWO
                        top, left, tw, th = top + top_off, left + left_off, sfac * w_, sfac * h_
                out.append(Nested(top, left, tw, th, content=self.layout_content(ch, 0), child_nodes_count=1))
        n.last_layout = (tiles, (sw, sh))
        return Nested(ZERO, ZERO, sw, sh, content=("color", rgba(c.background_color)),
                      child_nodes_count=sum(l.child_nodes_count for l in out), children=out)


# ---- transformations/layout/flatten.rs ------------------------------------------------------------------------------
def flatten(root: Nested, input_resolutions, res_w, res_h):
    shadow, layouts = inner_flatten(root, 0, [])
    out = []
    for l in shadow + layouts:
        if should_render(l, input_resolutions, res_w, res_h):
            out.append(fix_final(l))
    return out


def inner_flatten(n: Nested, offset, parent_masks):
    content = n.content
    if content[0] == "child":
        content = ("child", content[1] + offset, content[2], content[3])
        offset += 1
    layout = render_layout(n, content, parent_masks)
    shadows = [box_shadow_layout(n, sh, parent_masks) for sh in n.box_shadow]
    pm = list(parent_masks) + ([n.mask] if n.mask is not None else [])
    pm = child_parent_masks(n, pm)
    ch_sh, ch_l = [], []
    for ch in n.children:
        cnt = ch.child_nodes_count
        a, b = inner_flatten(ch, offset, list(pm))
        offset += cnt
        ch_sh += a
        ch_l += b
    ch_sh = [flatten_child(n, l) for l in ch_sh]
    ch_l = [flatten_child(n, l) for l in ch_l]
    return shadows, [layout] + ch_sh + ch_l


def render_layout(n, content, parent_masks):
    if content[0] == "color":
        return Render(n.top, n.left, n.width, n.height, n.rotation, n.border_radius, list(parent_masks), "color",
                      color=content[1], border_color=n.border_color, border_width=n.border_width)
    if content[0] == "child":
        return Render(n.top, n.left, n.width, n.height, n.rotation, n.border_radius, list(parent_masks), "child",
                      border_color=n.border_color, border_width=n.border_width, index=content[1],
                      crop=Crop(ZERO, ZERO, content[2], content[3]))
    return Render(n.top, n.left, n.width, n.height, n.rotation, n.border_radius, list(parent_masks), "color",
                  color=(0, 0, 0, 0), border_color=n.border_color, border_width=n.border_width)


def box_shadow_layout(n, sh, parent_masks):
    blur = F(sh.blur_radius)
    return Render(n.top + F(sh.offset_y), n.left + F(sh.offset_x), n.width, n.height, n.rotation,
                  n.border_radius.add(blur / TWO), list(parent_masks), "shadow", color=rgba(sh.color), blur_radius=blur)


def child_parent_masks(n, masks):
    k = fmin(n.scale_x, n.scale_y)
    return [Mask(m.radius.div(k), (m.top - n.top) / n.scale_y, (m.left - n.left) / n.scale_x, m.width / n.scale_x,
                 m.height / n.scale_y) for m in masks]


def parent_parent_masks(n, masks):
    k = fmin(n.scale_x, n.scale_y)
    return [Mask(m.radius.mul(k), (m.top * n.scale_y) + n.top, (m.left * n.scale_x) + n.left, m.width * n.scale_x,
                 m.height * n.scale_y) for m in masks]


def flatten_child(n, ch: Render):
    us = fmin(n.scale_x, n.scale_y)
    if n.crop is None:
        r = Render(n.top + (ch.top * n.scale_y), n.left + (ch.left * n.scale_x), ch.width * n.scale_x, ch.height * n.scale_y,
                   ch.rotation + n.rotation, ch.border_radius.mul(us), parent_parent_masks(n, ch.masks), ch.kind,
                   color=ch.color, border_color=ch.border_color, border_width=ch.border_width, blur_radius=ch.blur_radius,
                   index=ch.index, crop=ch.crop)
        if ch.kind == "shadow":
            r.blur_radius = ch.blur_radius * us
        else:
            r.border_width = ch.border_width * us
        return r
    raise NotImplementedError("NestedLayout.crop is never set by View / Rescaler / Tiles (scene/**: crop: None everywhere)")


def should_render(l: Render, input_resolutions, res_w, res_h):
    if l.width <= 0.0 or l.height <= 0.0 or l.top > F(res_h) or l.left > F(res_w):
        return False
    if l.kind == "color":
        if l.color[3] == 0:
            return l.border_color[3] != 0 or l.border_width > 0.0
        return True
    if l.kind == "child":
        size = input_resolutions[l.index] if 0 <= l.index < len(input_resolutions) else None
        if size is not None and (l.crop.left > F(size[0]) or l.crop.top > F(size[1])):
            return False
        if l.crop.top + l.crop.height < 0.0 or l.crop.left + l.crop.width < 0.0:
            return False
        return True
    return l.color[3] != 0


def fix_final(l: Render):
    if l.kind in ("color", "child") and l.border_width < 1.0:
        l.border_width = ZERO
    keep = []
    for m in l.masks:
        r = m.radius
        max_top, max_bottom = fmax(r.tl, r.tr), fmax(r.bl, r.br)
        max_left, max_right = fmax(r.tl, r.bl), fmax(r.tr, r.br)
        skip = (m.top + max_top <= l.top and m.left + max_left <= l.left and
                m.left + m.width - max_right >= l.left + l.width and m.top + m.height - max_bottom >= l.top + l.height)
        if not skip:
            keep.append(m)
    l.masks = keep
    return l


# ---- entry points ---------------------------------------------------------------------------------------------------
class StatefulScene:
    """One output of scene_state.rs: the scene copy (previous state of the next update) and the render-graph copy."""

    def __init__(self, out_w, out_h):
        self.out_w, self.out_h = out_w, out_h
        self.last_ns = 0
        self.resolutions = {}          # input resolutions of the last render
        self.scene_tree = None
        self.render_tree = None

    def update_scene(self, scene):
        if self.scene_tree is not None and self.scene_tree.kind != "input":   # recalculate_layout at last_pts
            Engine(self.last_ns).layout(self.scene_tree, F(self.out_w), F(self.out_h))
        prev = self.scene_tree.with_id({}) if self.scene_tree is not None else {}
        ctx = {"prev": prev, "last_ns": self.last_ns, "resolutions": dict(self.resolutions)}
        self.scene_tree = SNode(scene, ctx)
        self.render_tree = self.scene_tree.clone()

    def layouts(self, pts, resolutions_by_input_id):
        """render at pts (seconds): flattened RenderLayout list and the root node resolution"""
        pts_ns = to_ns(pts)
        self.last_ns, self.resolutions = pts_ns, dict(resolutions_by_input_id)   # register_render_event
        root = self.render_tree
        if root.kind == "input":
            return [], (0, 0)
        eng = Engine(pts_ns)
        leaves = root.node_children()
        in_res = [resolutions_by_input_id.get(k.comp.input_id) for k in leaves]
        eng.update_state(root, in_res)
        p = eng.position(root)           # SizedLayoutComponent::resolution; Size -> Resolution truncates
        w = p[1] if p[1] is not None else F(self.out_w)
        h = p[2] if p[2] is not None else F(self.out_h)
        rw, rh = int(np.trunc(w)), int(np.trunc(h))
        nested = eng.layout(root, F(self.out_w), F(self.out_h))
        return flatten(nested, in_res, rw, rh), (rw, rh)


def layouts(scene, out_w, out_h, resolutions_by_input_id, pts=0.0):
    """a freshly registered scene rendered once"""
    st = StatefulScene(out_w, out_h)
    st.update_scene(scene)
    return st.layouts(pts, resolutions_by_input_id)
