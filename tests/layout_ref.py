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


# ---- stateless view of a component at one pts -------------------------------------------------------------------
def is_layout(c):
    return isinstance(c, (s.ViewComponent, s.RescalerComponent, s.TilesComponent))


def node_children(c):                # layout.rs:84-93
    out = []
    for ch in children_of(c):
        out += node_children(ch) if is_layout(ch) else [ch]
    return out


def children_of(c):
    if isinstance(c, s.RescalerComponent):
        return [c.child if c.child is not None else s.ViewComponent()]
    return list(getattr(c, "children", []))


class Engine:
    """One evaluation at a fixed pts.  `pos_of(component)` answers the (possibly interpolated) position / size state of
    View and Rescaler; `tiles_of(component, size)` the (possibly interpolated) tile list of Tiles."""

    def __init__(self, input_sizes, view_state=None, tiles_state=None):
        self.input_sizes = input_sizes          # id(InputStreamComponent) -> (w, h) or None
        self.view_state = view_state or (lambda c: None)
        self.tiles_state = tiles_state or (lambda c, size, end: end)

    # -- position / sizes ----------------------------------------------------------------------------------------
    def params(self, c):
        """View / Rescaler parameters at this pts (transition-interpolated when the stateful scene says so)"""
        st = self.view_state(c)
        if st is not None:
            return st
        return base_params(c)

    def position(self, c):           # Position of a layout component
        if isinstance(c, s.TilesComponent):
            return ("static", None if c.width is None else F(c.width), None if c.height is None else F(c.height))
        return self.params(c)["position"]

    def width(self, c):              # scene.rs:102-114
        if isinstance(c, s.InputStreamComponent):
            return self.size_of_input(c)[0]
        p = self.position(c)
        return p[1]

    def height(self, c):
        if isinstance(c, s.InputStreamComponent):
            return self.size_of_input(c)[1]
        p = self.position(c)
        return p[2]

    def size_of_input(self, c):      # layout.rs:99-108: missing -> 0 x 0
        r = self.input_sizes.get(id(c))
        return (F(r[0]), F(r[1])) if r is not None else (ZERO, ZERO)

    def layout_content(self, c, index):   # layout.rs:133-157
        if is_layout(c):
            return ("none",)
        w, h = self.size_of_input(c)
        return ("child", index, w, h)

    # -- dispatch ------------------------------------------------------------------------------------------------
    def layout(self, c, w, h):
        if isinstance(c, s.ViewComponent):
            return self.view_layout(c, F(w), F(h))
        if isinstance(c, s.RescalerComponent):
            return self.rescaler_layout(c, F(w), F(h))
        return self.tiles_layout(c, F(w), F(h))

    # -- layout.rs:159-236 -----------------------------------------------------------------------------------------
    def absolute_child(self, child, pos, pw, ph):
        _, pwid, phei, hor, ver, rot = pos
        width = pwid if pwid is not None else pw
        height = phei if phei is not None else ph
        top = ver[1] if ver[0] == "top" else (ph - ver[1]) - height
        left = hor[1] if hor[0] == "left" else (pw - hor[1]) - width
        content = self.layout_content(child, 0)
        if is_layout(child):
            cl = self.layout(child, width, height)
            cnt = cl.child_nodes_count + (1 if content[0] == "child" else 0)
            return Nested(top, left, width, height, rotation=rot, content=content, child_nodes_count=cnt, children=[cl])
        return Nested(top, left, width, height, rotation=rot, content=content, child_nodes_count=1 if content[0] == "child" else 0)

    # -- view_component/layout.rs ---------------------------------------------------------------------------------
    def view_layout(self, c, sw, sh):
        P = self.params(c)
        bw = P["border_width"]
        cw = fmax(sw - TWO * bw, ZERO)
        chh = fmax(sh - TWO * bw, ZERO)
        br = P["border_radius"].clip_to_size(sw, sh)
        kids = children_of(c)
        static_child_size = self.static_child_size(c, P, cw, chh, kids)
        ov = c.overflow
        if ov == s.Overflow.Visible:
            scale, mask = ONE, None
        else:
            scale = ONE if ov == s.Overflow.Hidden else self.scale_factor_for_overflow_fit(c, cw, chh, kids)
            mask = Mask(br.sub(bw), bw, bw, cw, chh)
        static_offset = bw / scale
        pad = c.padding
        out = []
        for ch in kids:
            if is_layout(ch):
                pos = self.position(ch)
            else:
                pos = ("static", self.width(ch), self.height(ch))
            if pos[0] == "static":
                _, w_, h_ = pos
                pbw = bw / scale
                if c.direction == s.ViewChildrenDirection.Row:
                    width = w_ if w_ is not None else static_child_size
                    height = h_ if h_ is not None else chh - (F(pad.top) + F(pad.bottom))
                    top = pbw + F(pad.top)
                    left = static_offset + F(pad.left)
                    static_offset = static_offset + width
                else:
                    height = h_ if h_ is not None else static_child_size
                    width = w_ if w_ is not None else cw - (F(pad.left) + F(pad.right))
                    top = static_offset + F(pad.top)
                    left = pbw + F(pad.left)
                    static_offset = static_offset + height
                if is_layout(ch):
                    cl = self.layout(ch, width, height)
                    out.append(Nested(top, left, width, height, content=("none",), child_nodes_count=cl.child_nodes_count, children=[cl]))
                else:
                    out.append(Nested(top, left, width, height, content=self.layout_content(ch, 0), child_nodes_count=1))
            else:
                out.append(self.absolute_child(ch, pos, sw, sh))
        return Nested(ZERO, ZERO, sw, sh, scale_x=scale, scale_y=scale, mask=mask, content=("color", rgba(c.background_color)),
                      child_nodes_count=sum(l.child_nodes_count for l in out), children=out, border_width=bw,
                      border_color=rgba(c.border_color), border_radius=br, box_shadow=list(c.box_shadow))

    def static_children(self, kids):
        return [k for k in kids if not is_layout(k) or self.position(k)[0] == "static"]

    def sum_static_children_sizes(self, c, kids):
        acc = ZERO
        for k in self.static_children(kids):
            v = self.width(k) if c.direction == s.ViewChildrenDirection.Row else self.height(k)
            acc = acc + (v if v is not None else ZERO)
        return acc

    def static_child_size(self, c, P, cw, chh, kids):
        pad = c.padding
        row = c.direction == s.ViewChildrenDirection.Row
        max_size = cw - (F(pad.left) + F(pad.right)) if row else chh - (F(pad.top) + F(pad.bottom))
        unknown = sum(1 for k in self.static_children(kids) if (self.width(k) if row else self.height(k)) is None)
        total = self.sum_static_children_sizes(c, kids)
        if unknown == 0:
            return ZERO
        return fmax(ZERO, (max_size - total) / F(unknown))

    def scale_factor_for_overflow_fit(self, c, cw, chh, kids):
        row = c.direction == s.ViewChildrenDirection.Row
        sum_size = fmax(self.sum_static_children_sizes(c, kids), F(0.000000001))
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
    def rescaler_layout(self, c, sw, sh):
        P = self.params(c)
        bw = P["border_width"]
        cw = fmax(sw - (TWO * bw), ZERO)
        chh = fmax(sh - (TWO * bw), ZERO)
        child = children_of(c)[0]
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
        if is_layout(child):
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
                      border_radius=br, box_shadow=list(c.box_shadow))

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

    def end_tiles(self, c, lw, lh):
        kids = children_of(c)
        n = len(kids)
        if n == 0:
            return []
        best = (1, n)
        best_w = ZERO
        for rows in range(1, n + 1):
            cols = -(-n // rows)
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
            in_row = cols if row < rows - 1 else n - (rows - 1) * cols
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
            cid = getattr(k, "id", None)
            if cid is not None:
                tid = ("id", cid)
            else:
                tid = ("index", idx)
                idx += 1
            tiles.append({"id": tid, "top": t[0], "left": t[1], "width": t[2], "height": t[3]})
        return tiles

    def tiles_layout(self, c, sw, sh):
        end = self.end_tiles(c, sw, sh)
        tiles = self.tiles_state(c, (sw, sh), end)
        out = []
        for ch, tile in zip(children_of(c), tiles):
            if tile is None:     # child_nodes_placeholder
                cnt = len(node_children(ch)) if is_layout(ch) else 1
                out.append(Nested(ZERO, ZERO, ZERO, ZERO, content=("none",), child_nodes_count=cnt))
                continue
            if is_layout(ch):
                cl = self.layout(ch, tile["width"], tile["height"])
                out.append(Nested(tile["top"], tile["left"], tile["width"], tile["height"], content=("none",),
                                  child_nodes_count=cl.child_nodes_count, children=[cl]))
            else:
                w_, h_ = self.width(ch), self.height(ch)
                top, left, tw, th = tile["top"], tile["left"], tile["width"], tile["height"]
                if w_ is not None and h_ is not None:   # fit_into_tile
                    sfac = fmin(tw / w_, th / h_)
                    top_off = (th - sfac * h_) / TWO
                    left_off = (tw - sfac * w_) / TWO
                    top, left, tw, th = top + top_off, left + left_off, sfac * w_, sfac * h_
                out.append(Nested(top, left, tw, th, content=self.layout_content(ch, 0), child_nodes_count=1))
        return Nested(ZERO, ZERO, sw, sh, content=("color", rgba(c.background_color)),
                      child_nodes_count=sum(l.child_nodes_count for l in out), children=out)


def base_params(c):
    """View / Rescaler parameters that transitions interpolate (view_component.rs / rescaler_component.rs)"""
    p = c.position
    if p.absolute:
        hor = ("right", F(p.right)) if p.right is not None else ("left", F(p.left or 0.0))
        ver = ("bottom", F(p.bottom)) if p.bottom is not None else ("top", F(p.top or 0.0))
        pos = ("absolute", None if p.width is None else F(p.width), None if p.height is None else F(p.height), hor, ver,
               F(p.rotation_degrees))
    else:
        pos = ("static", None if p.width is None else F(p.width), None if p.height is None else F(p.height))
    return {"position": pos, "border_width": F(c.border_width), "border_radius": Radius.of(c.border_radius)}


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
def leaf_inputs(c):
    return [k for k in node_children(c)] if is_layout(c) else [c]


def root_resolution(engine, scene, out_w, out_h):
    """SizedLayoutComponent::resolution (layout.rs:243-256): Size -> Resolution truncates (types/convert.rs:14-21)"""
    p = engine.position(scene)
    w = p[1] if p[1] is not None else F(out_w)
    h = p[2] if p[2] is not None else F(out_h)
    return int(np.trunc(w)), int(np.trunc(h))


def layouts(scene, out_w, out_h, resolutions_by_input_id, engine_factory=None):
    """Flattened RenderLayout list of `scene` registered on an out_w x out_h output.  `resolutions_by_input_id`:
    {input_id: (w, h)} of the inputs that have a (fresh) frame.  Returns (layouts, (root_w, root_h))."""
    leaves = leaf_inputs(scene)
    sizes = {id(k): resolutions_by_input_id.get(k.input_id) for k in leaves}
    eng = engine_factory(sizes) if engine_factory else Engine(sizes)
    nested = eng.layout(scene, F(out_w), F(out_h))
    rw, rh = root_resolution(eng, scene, out_w, out_h)
    in_res = [resolutions_by_input_id.get(k.input_id) for k in leaves]
    return flatten(nested, in_res, rw, rh), (rw, rh)
