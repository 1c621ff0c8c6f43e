// scene.h -- host-side scene model of the compositor (C++ restatement of smelter-render/src/scene/*).
//
// The reference keeps this math on the CPU too: Component tree -> stateful tree
// (transitions) -> NestedLayout -> flatten -> RenderLayout[].  Only the resulting RenderLayout list
// reaches the GPU.  All arithmetic is f32 (f64 for easing) in the reference's order of operations.
#pragma once

#include <cstddef>
#include <cstdint>
#include <map>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "../../include/smelter_b200.h"

namespace smr {

struct Size {
    float width = 0.0f, height = 0.0f;
};

struct Resolution {
    size_t width = 0, height = 0;
    bool operator==(const Resolution &o) const { return width == o.width && height == o.height; }
};

struct RGBA {
    uint8_t r = 0, g = 0, b = 0, a = 0;
    bool operator==(const RGBA &o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
};

// scene/types.rs:91-160
struct BorderRadius {
    float top_left = 0, top_right = 0, bottom_right = 0, bottom_left = 0;
    BorderRadius clip_to_size(Size size) const;
    BorderRadius operator*(float rhs) const;
    BorderRadius operator/(float rhs) const { return *this * (1.0f / rhs); }
    BorderRadius operator+(float rhs) const;
    BorderRadius operator-(float rhs) const { return *this + (-rhs); }
    bool operator==(const BorderRadius &o) const {
        return top_left == o.top_left && top_right == o.top_right && bottom_right == o.bottom_right &&
               bottom_left == o.bottom_left;
    }
};

struct BoxShadow {
    float offset_x = 0, offset_y = 0, blur_radius = 0;
    RGBA color;
    bool operator==(const BoxShadow &o) const {
        return offset_x == o.offset_x && offset_y == o.offset_y && blur_radius == o.blur_radius && color == o.color;
    }
};

struct Padding {
    float top = 0, right = 0, bottom = 0, left = 0;
    float horizontal() const { return left + right; }
    float vertical() const { return top + bottom; }
    bool operator==(const Padding &o) const {
        return top == o.top && right == o.right && bottom == o.bottom && left == o.left;
    }
};

using OptF = std::optional<float>;

// components.rs:199-206 + types.rs:66-86
struct Position {
    bool absolute = false;
    OptF width, height;
    bool from_right = false;
    float horizontal_offset = 0;
    bool from_bottom = false;
    float vertical_offset = 0;
    float rotation_degrees = 0;
    Position with_border(float border_width) const;   // components/position.rs:6-29
    Position with_padding(const Padding &p) const;    // components/position.rs:31-54
    bool operator==(const Position &o) const;
};

struct InterpolationKind {
    int kind = SMR_INTERP_LINEAR;
    double x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    double state(double t) const;  // transition.rs:108-118
};

struct Transition {
    uint64_t duration_ns = 0;
    InterpolationKind interpolation;
    bool should_interrupt = false;
};

// scene::Component (scene.rs:50-60); only the variants on the compositor path
struct Component {
    int type = SMR_COMPONENT_VIEW;
    std::optional<std::string> id;
    std::vector<Component> children;
    std::string input_id;
    Position position;
    std::optional<Transition> transition;
    BorderRadius border_radius;
    float border_width = 0;
    RGBA border_color;
    std::vector<BoxShadow> box_shadow;
    int direction = SMR_DIRECTION_ROW;
    int overflow = SMR_OVERFLOW_HIDDEN;
    RGBA background_color;
    Padding padding;
    int rescale_mode = SMR_RESCALE_FIT;
    int horizontal_align = SMR_HALIGN_CENTER;
    int vertical_align = SMR_VALIGN_CENTER;
    OptF tiles_width, tiles_height;
    uint32_t tile_aspect_w = 16, tile_aspect_h = 9;
    float tiles_margin = 0, tiles_padding = 0;
};

// Converts the C tree; returns false + message for variants outside the hot path.
bool component_from_c(const smr_component *c, Component &out, std::string &err, int depth = 0);

// ----- transformations/layout.rs:40-165 -------------------------------------------------------
struct Crop {
    float top = 0, left = 0, width = 0, height = 0;
};

struct Mask {
    BorderRadius radius;
    float top = 0, left = 0, width = 0, height = 0;
};

struct LayoutContent {
    enum Kind { Color, ChildNode, None } kind = None;
    RGBA color;
    size_t index = 0;
    Size size;
};

struct RenderLayout {
    enum Kind { Color = 1, ChildNode = 0, BoxShadow = 2 };  // numeric = layout_type of the shader
    float top = 0, left = 0, width = 0, height = 0, rotation_degrees = 0;
    BorderRadius border_radius;
    std::vector<Mask> masks;
    Kind kind = Color;
    RGBA color;         // Color / BoxShadow
    RGBA border_color;  // Color / ChildNode
    float border_width = 0;
    size_t index = 0;   // ChildNode
    Crop crop;          // ChildNode
    float blur_radius = 0;
};

struct NestedLayout {
    float top = 0, left = 0, width = 0, height = 0, rotation_degrees = 0;
    float scale_x = 1.0f, scale_y = 1.0f;
    std::optional<Crop> crop;
    std::optional<Mask> mask;
    LayoutContent content;
    float border_width = 0;
    RGBA border_color;
    BorderRadius border_radius;
    std::vector<BoxShadow> box_shadow;
    std::vector<NestedLayout> children;
    size_t child_nodes_count = 0;

    static NestedLayout child_nodes_placeholder(size_t child_nodes_count);  // layout.rs:280-304
    // layout/flatten.rs:10-22
    std::vector<RenderLayout> flatten(const std::vector<std::optional<Resolution>> &input_resolutions,
                                      Resolution resolution) const;
};

// ----- stateful components (scene/{view,rescaler,tiles}_component.rs) --------------------------
struct ViewParam {
    std::optional<std::string> id;
    int direction = SMR_DIRECTION_ROW;
    Position position;
    int overflow = SMR_OVERFLOW_HIDDEN;
    RGBA background_color;
    BorderRadius border_radius;
    float border_width = 0;
    RGBA border_color;
    std::vector<BoxShadow> box_shadow;
    Padding padding;
    bool operator==(const ViewParam &o) const;
};

struct RescalerParam {
    std::optional<std::string> id;
    Position position;
    int mode = SMR_RESCALE_FIT;
    int horizontal_align = SMR_HALIGN_CENTER;
    int vertical_align = SMR_VALIGN_CENTER;
    BorderRadius border_radius;
    float border_width = 0;
    RGBA border_color;
    std::vector<BoxShadow> box_shadow;
    bool operator==(const RescalerParam &o) const;
};

struct TilesParam {
    std::optional<std::string> id;
    OptF width, height;
    RGBA background_color;
    uint32_t aspect_w = 16, aspect_h = 9;
    float margin = 0, padding = 0;
    int horizontal_align = SMR_HALIGN_CENTER;
    int vertical_align = SMR_VALIGN_CENTER;
    bool operator==(const TilesParam &o) const;
};

struct TileId {
    bool is_component = false;
    std::string component_id;
    size_t index = 0;
    bool operator==(const TileId &o) const {
        return is_component == o.is_component && component_id == o.component_id && index == o.index;
    }
};

struct Tile {
    TileId id;
    float top = 0, left = 0, width = 0, height = 0;
};
using OptTile = std::optional<Tile>;
using TilesSnapshot = std::pair<std::vector<OptTile>, Size>;

// scene/transition.rs:20-106
struct TransitionState {
    double offset_progress = 0.0, offset_state = 0.0;
    uint64_t start_pts_ns = 0;
    uint64_t duration_ns = 0;
    InterpolationKind interpolation;

    static std::optional<TransitionState> create(const std::optional<Transition> &current,
                                                 const std::optional<TransitionState> &previous,
                                                 bool props_changed, bool interrupt_previous,
                                                 uint64_t last_pts_ns);
    double state(uint64_t pts_ns) const;
    bool is_finished(uint64_t pts_ns) const { return start_pts_ns + duration_ns <= pts_ns; }
};

struct Stateful {
    enum Kind { InputStream, View, Tiles, Rescaler } kind = View;
    // InputStream (scene/input_stream_component.rs)
    std::string input_id;
    std::optional<std::string> input_component_id;
    Size size;
    // View
    std::optional<ViewParam> view_start;
    ViewParam view_end;
    // Rescaler
    std::optional<RescalerParam> rescaler_start;
    RescalerParam rescaler_end;
    // Tiles
    TilesParam tiles;
    std::optional<TilesSnapshot> tiles_start, tiles_last_layout;

    std::optional<TransitionState> transition;
    std::vector<Stateful> children;  // Rescaler: exactly one

    bool is_layout() const { return kind != InputStream; }
    const std::optional<std::string> &component_id() const;
    OptF width(uint64_t pts) const;   // scene.rs:105-117
    OptF height(uint64_t pts) const;  // scene.rs:119-131
    Position position(uint64_t pts) const;
    NestedLayout layout(Size size, uint64_t pts);  // scene/layout.rs:44-50
    void node_children(std::vector<const Stateful *> &out) const;  // scene/layout.rs:95-103
    size_t node_children_count() const;
    void update_state(const std::optional<Resolution> *inputs, size_t n);  // scene/layout.rs:105-137
};

// scene/scene_state.rs
struct OutputNode {
    bool root_is_input = false;
    std::string root_input_id;
    Stateful layout_root;                     // LayoutNode.root.component (the render graph's clone)
    Size size;                                // SizedLayoutComponent.size
    std::vector<std::string> child_input_ids; // node children, DFS order
    Resolution resolution;

    // scene::LayoutNode as LayoutProvider (scene/layout.rs:31-41, 240-261)
    Resolution layout_resolution(uint64_t pts) const;
    NestedLayout layouts(uint64_t pts, const std::vector<std::optional<Resolution>> &inputs);
};

class SceneState {
  public:
    void register_render_event(uint64_t pts_ns, std::map<std::string, Resolution> input_resolutions);
    void unregister_output(const std::string &output_id);
    // returns false and fills err on SceneError
    bool update_scene(const std::string &output_id, const Component &root, Resolution resolution,
                      OutputNode &out, std::string &err);
    uint64_t last_pts() const { return last_pts_ns_; }

  private:
    struct OutputSceneState {
        Stateful root;
        Resolution resolution;
    };
    std::map<std::string, Component> output_scenes_;
    std::map<std::string, OutputSceneState> output_states_;
    uint64_t last_pts_ns_ = 0;
    std::map<std::string, Resolution> input_resolutions_;
};

double cubic_bezier_easing(double progress, double x1, double y1, double x2, double y2);
double bounce_easing(double t);

}  // namespace smr
