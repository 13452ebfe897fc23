// fluxhost.cpp — see fluxhost.hpp.  A small YAML reader for the subset serde_yaml sees in scenes/*.yml
// (block maps and sequences, flow sequences/maps, anchors & aliases, comments, plain/quoted scalars), the
// SceneData model with serde's "every field required, unknown keys ignored" behaviour, flattening for the
// C-ABI, and the GpuWorker that drives libfluxb200.so.
#include "fluxhost.hpp"
#include "node.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <thread>

namespace flux {
namespace {

// ------------------------------------------------------------------------------------------------
// YAML subset
// ------------------------------------------------------------------------------------------------
using detail::Node;

struct Line {
    int indent;
    std::string text;   // without indentation, comments and trailing blanks
    int number;
};

std::string rtrim(std::string s) {
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back();
    return s;
}
std::string ltrim(const std::string &s) {
    size_t i = 0;
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++;
    return s.substr(i);
}
std::string trim(const std::string &s) { return rtrim(ltrim(s)); }

// strip a trailing comment: '#' at line start or preceded by a blank, outside quotes
std::string strip_comment(const std::string &s) {
    char q = 0;
    for (size_t i = 0; i < s.size(); i++) {
        const char c = s[i];
        if (q) {
            if (c == q) q = 0;
        } else if (c == '"' || c == '\'') {
            q = c;
        } else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) {
            return s.substr(0, i);
        }
    }
    return s;
}

class Parser {
  public:
    explicit Parser(const std::string &text) {
        std::istringstream in(text);
        std::string raw;
        int no = 0;
        while (std::getline(in, raw)) {
            no++;
            std::string s = rtrim(strip_comment(raw));
            if (s.empty()) continue;
            if (s == "---" || s == "...") continue;
            int ind = 0;
            while ((size_t)ind < s.size() && s[ind] == ' ') ind++;
            if ((size_t)ind < s.size() && s[ind] == '\t') fail(no, "tabs are not allowed for indentation");
            lines_.push_back(Line{ind, s.substr(ind), no});
        }
    }

    Node parse() {
        if (lines_.empty()) return Node{};
        pos_ = 0;
        Node n = block(lines_[0].indent);
        if (pos_ != lines_.size()) fail(lines_[pos_].number, "unexpected indentation");
        return n;
    }

  private:
    std::vector<Line> lines_;
    size_t pos_ = 0;
    std::map<std::string, Node> anchors_;

    [[noreturn]] static void fail(int line, const std::string &msg) {
        throw Error("yaml: line " + std::to_string(line) + ": " + msg);
    }
    static bool is_seq_item(const std::string &t) { return t == "-" || (t.size() > 1 && t[0] == '-' && t[1] == ' '); }

    // a block collection whose entries start at column `indent`
    Node block(int indent) {
        if (pos_ >= lines_.size()) return Node{};
        return is_seq_item(lines_[pos_].text) ? sequence(indent) : mapping(indent);
    }

    Node sequence(int indent) {
        Node n;
        n.kind = Node::Seq;
        while (pos_ < lines_.size() && lines_[pos_].indent == indent && is_seq_item(lines_[pos_].text)) {
            Line &ln = lines_[pos_];
            std::string rest = ln.text.size() > 1 ? ltrim(ln.text.substr(1)) : "";
            if (rest.empty()) {   // "-" alone: the item is the deeper block
                pos_++;
                if (pos_ < lines_.size() && lines_[pos_].indent > indent) n.seq.push_back(block(lines_[pos_].indent));
                else n.seq.push_back(Node{});
                continue;
            }
            const int col = indent + (int)(ln.text.size() - rest.size());
            if (looks_like_key(rest)) {   // "- key: value": a mapping whose first key sits on this line
                ln.indent = col;
                ln.text = rest;
                n.seq.push_back(mapping(col));
            } else {
                pos_++;
                n.seq.push_back(value(rest, ln.number, indent));
            }
        }
        if (pos_ < lines_.size() && lines_[pos_].indent > indent) fail(lines_[pos_].number, "unexpected indentation in sequence");
        return n;
    }

    // "key:" or "key: value" with a plain (unquoted, non-flow) key
    static bool looks_like_key(const std::string &t) {
        if (t.empty() || t[0] == '[' || t[0] == '{' || t[0] == '"' || t[0] == '\'' || t[0] == '&' || t[0] == '*') return false;
        const size_t c = key_colon(t);
        return c != std::string::npos;
    }
    static size_t key_colon(const std::string &t) {
        for (size_t i = 0; i < t.size(); i++) {
            if (t[i] == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
            if (t[i] == '[' || t[i] == '{' || t[i] == '"' || t[i] == '\'') return std::string::npos;
        }
        return std::string::npos;
    }

    Node mapping(int indent) {
        Node n;
        n.kind = Node::Map;
        while (pos_ < lines_.size() && lines_[pos_].indent == indent && !is_seq_item(lines_[pos_].text)) {
            const Line ln = lines_[pos_];
            const size_t c = key_colon(ln.text);
            if (c == std::string::npos) fail(ln.number, "expected `key: value`");
            const std::string key = trim(ln.text.substr(0, c));
            const std::string rest = trim(ln.text.substr(c + 1));
            pos_++;
            Node v = value(rest, ln.number, indent);
            bool replaced = false;
            for (auto &kv : n.map)
                if (kv.first == key) {
                    kv.second = v;
                    replaced = true;
                }
            if (!replaced) n.map.emplace_back(key, std::move(v));
        }
        if (pos_ < lines_.size() && lines_[pos_].indent > indent) fail(lines_[pos_].number, "unexpected indentation in mapping");
        return n;
    }

    // the value after "key:" / "- ": inline text and/or the following deeper block
    Node value(std::string rest, int line_no, int parent_indent) {
        std::string anchor;
        if (!rest.empty() && rest[0] == '&') {
            size_t e = rest.find(' ');
            anchor = rest.substr(1, e == std::string::npos ? std::string::npos : e - 1);
            rest = e == std::string::npos ? "" : trim(rest.substr(e + 1));
        }
        Node v;
        if (rest.empty()) {
            if (pos_ < lines_.size() && lines_[pos_].indent > parent_indent) v = block(lines_[pos_].indent);
            else if (pos_ < lines_.size() && lines_[pos_].indent == parent_indent && is_seq_item(lines_[pos_].text))
                v = sequence(parent_indent);   // "key:\n- item" (sequence at the key's own column)
        } else if (rest[0] == '*') {
            auto it = anchors_.find(trim(rest.substr(1)));
            if (it == anchors_.end()) fail(line_no, "unknown alias `" + rest + "`");
            v = it->second;
        } else if (rest[0] == '[' || rest[0] == '{') {
            // a flow collection may continue on the following lines until the brackets balance
            std::string flow = rest;
            while (!balanced(flow)) {
                if (pos_ >= lines_.size()) fail(line_no, "unterminated flow collection");
                flow += " " + lines_[pos_].text;
                pos_++;
            }
            size_t p = 0;
            v = flow_value(flow, p, line_no);
            if (trim(flow.substr(p)) != "") fail(line_no, "trailing characters after flow collection");
        } else {
            v = scalar(rest);
        }
        if (!anchor.empty()) anchors_[anchor] = v;
        return v;
    }

    static bool balanced(const std::string &s) {
        int depth = 0;
        char q = 0;
        for (char c : s) {
            if (q) {
                if (c == q) q = 0;
            } else if (c == '"' || c == '\'') q = c;
            else if (c == '[' || c == '{') depth++;
            else if (c == ']' || c == '}') depth--;
        }
        return depth <= 0;
    }

    static Node scalar(std::string t) {
        Node n;
        n.kind = Node::Scalar;
        t = trim(t);
        if (t.size() >= 2 && (t.front() == '"' || t.front() == '\'') && t.back() == t.front()) {
            n.quoted = true;
            t = t.substr(1, t.size() - 2);
        } else if (t == "~" || t == "null") {
            n.kind = Node::Null;
        }
        n.scalar = t;
        return n;
    }

    Node flow_value(const std::string &s, size_t &p, int line_no) {
        while (p < s.size() && s[p] == ' ') p++;
        if (p >= s.size()) fail(line_no, "unexpected end of flow collection");
        if (s[p] == '[') {
            Node n;
            n.kind = Node::Seq;
            p++;
            for (;;) {
                while (p < s.size() && s[p] == ' ') p++;
                if (p < s.size() && s[p] == ']') { p++; break; }
                n.seq.push_back(flow_value(s, p, line_no));
                while (p < s.size() && s[p] == ' ') p++;
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; break; }
                fail(line_no, "expected `,` or `]` in flow sequence");
            }
            return n;
        }
        if (s[p] == '{') {
            Node n;
            n.kind = Node::Map;
            p++;
            for (;;) {
                while (p < s.size() && s[p] == ' ') p++;
                if (p < s.size() && s[p] == '}') { p++; break; }
                size_t c = s.find(':', p);
                if (c == std::string::npos) fail(line_no, "expected `key: value` in flow mapping");
                std::string key = trim(s.substr(p, c - p));
                p = c + 1;
                n.map.emplace_back(key, flow_value(s, p, line_no));
                while (p < s.size() && s[p] == ' ') p++;
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; break; }
                fail(line_no, "expected `,` or `}` in flow mapping");
            }
            return n;
        }
        if (s[p] == '*') {
            size_t e = s.find_first_of(",]} ", p);
            auto it = anchors_.find(s.substr(p + 1, e == std::string::npos ? std::string::npos : e - p - 1));
            if (it == anchors_.end()) fail(line_no, "unknown alias in flow collection");
            p = e == std::string::npos ? s.size() : e;
            return it->second;
        }
        size_t e = p;
        if (s[p] == '"' || s[p] == '\'') {
            e = s.find(s[p], p + 1);
            if (e == std::string::npos) fail(line_no, "unterminated string");
            e++;
        } else {
            e = s.find_first_of(",]}", p);
            if (e == std::string::npos) e = s.size();
        }
        Node n = scalar(s.substr(p, e - p));
        p = e;
        return n;
    }
};

// ------------------------------------------------------------------------------------------------
// serde-like field access
// ------------------------------------------------------------------------------------------------
const Node &req(const Node &m, const char *key, const char *what) {
    const Node *n = m.find(key);
    if (!n) throw Error(std::string(what) + ": missing field `" + key + "`");
    return *n;
}

double as_f64(const Node &n, const std::string &what) {
    if (n.kind == Node::Scalar && n.bin == Node::F64) return n.f;
    if (n.kind == Node::Scalar && n.bin == Node::U64) return (double)n.u;   // serde's f64 visitor takes integers
    if (n.kind == Node::Scalar && n.bin == Node::I64) return (double)(int64_t)n.u;
    if (n.kind == Node::Scalar && n.bin == Node::Text && !n.quoted) {
        const std::string &t = n.scalar;
        if (t == ".inf" || t == "+.inf" || t == ".Inf") return INFINITY;
        if (t == "-.inf" || t == "-.Inf") return -INFINITY;
        if (t == ".nan" || t == ".NaN") return NAN;
        char *end = nullptr;
        const double v = std::strtod(t.c_str(), &end);
        // YAML floats/ints only: no hex floats, no "inf"/"nan" spellings of strtod
        const bool plain = !t.empty() && t.find_first_not_of("+-0123456789.eE_") == std::string::npos;
        if (plain && end && *end == '\0') return v;
    }
    throw Error(what + ": invalid type: expected f64" + (n.kind == Node::Scalar ? ", got `" + n.scalar + "`" : ""));
}

uint32_t as_usize(const Node &n, const std::string &what) {
    if (n.kind == Node::Scalar && n.bin == Node::U64 && n.u <= 0xFFFFFFFFull) return (uint32_t)n.u;
    if (n.kind == Node::Scalar && n.bin == Node::Text && !n.quoted && !n.scalar.empty() && n.scalar.find_first_not_of("0123456789") == std::string::npos) {
        const unsigned long long v = std::strtoull(n.scalar.c_str(), nullptr, 10);
        if (v <= 0xFFFFFFFFull) return (uint32_t)v;
    }
    throw Error(what + ": invalid type: expected an unsigned integer" + (n.kind == Node::Scalar ? ", got `" + n.scalar + "`" : ""));
}

bool as_bool(const Node &n, const std::string &what) {
    if (n.kind == Node::Scalar && n.bin == Node::Bool) return n.u != 0;
    if (n.kind == Node::Scalar && n.bin == Node::Text && !n.quoted) {
        if (n.scalar == "true") return true;
        if (n.scalar == "false") return false;
    }
    throw Error(what + ": invalid type: expected a boolean" + (n.kind == Node::Scalar ? ", got `" + n.scalar + "`" : ""));
}

// Vector3 / Point3 / Color: a sequence of 3 numbers (serde also accepts a map for Color {r, g, b})
Vec3 as_vec3(const Node &n, const std::string &what) {
    if (n.kind == Node::Map) {
        return Vec3{as_f64(req(n, "r", what.c_str()), what), as_f64(req(n, "g", what.c_str()), what), as_f64(req(n, "b", what.c_str()), what)};
    }
    if (n.kind != Node::Seq || n.seq.size() != 3) throw Error(what + ": expected a sequence of 3 numbers");
    return Vec3{as_f64(n.seq[0], what), as_f64(n.seq[1], what), as_f64(n.seq[2], what)};
}

// externally tagged enum: a single-key map { Variant: {...} }; from CBOR also serde_cbor's array form
// [ "Variant", {...} ] (the only form serde_cbor < 0.10 writes; the reference pins 0.9.0)
struct VariantRef {
    const std::string &first;
    const Node &second;
};
VariantRef variant_of(const Node &n, const char *what) {
    if (n.kind == Node::Map && n.map.size() == 1) return VariantRef{n.map[0].first, n.map[0].second};
    if (n.cbor && n.kind == Node::Seq && n.seq.size() == 2 && n.seq[0].kind == Node::Scalar && n.seq[0].bin == Node::Text)
        return VariantRef{n.seq[0].scalar, n.seq[1]};
    throw Error(std::string(what) + ": expected a single-key map (externally tagged enum)");
}

MaterialData material_from_yaml(const Node &n) {
    const auto &kv = variant_of(n, "material");
    const std::string &tag = kv.first;
    const Node &m = kv.second;
    const char *w = tag.c_str();
    if (tag == "Matte")
        return MatteData{as_vec3(req(m, "diffuse_color", w), "diffuse_color"), as_vec3(req(m, "ambient_color", w), "ambient_color"),
                         as_f64(req(m, "diffuse_coefficient", w), "diffuse_coefficient")};
    if (tag == "Emissive") return EmissiveData{as_vec3(req(m, "color", w), "color"), as_f64(req(m, "power", w), "power")};
    if (tag == "Reflective")
        return ReflectiveData{as_f64(req(m, "reflect_amount", w), "reflect_amount"), as_vec3(req(m, "reflect_color", w), "reflect_color")};
    if (tag == "GlossyReflective")
        return GlossyReflectiveData{as_f64(req(m, "reflect_amount", w), "reflect_amount"),
                                    as_vec3(req(m, "reflect_color", w), "reflect_color"),
                                    as_f64(req(m, "reflect_exponent", w), "reflect_exponent")};
    throw Error("unknown variant `" + tag + "`, expected one of `Matte`, `Emissive`, `Reflective`, `GlossyReflective`");
}

ShapeData shape_from_yaml(const Node &n) {
    const auto &kv = variant_of(n, "shape");
    const std::string &tag = kv.first;
    const Node &s = kv.second;
    const char *w = tag.c_str();
    if (tag == "Sphere")
        return SphereData{as_vec3(req(s, "center", w), "center"), as_f64(req(s, "radius", w), "radius"),
                          material_from_yaml(req(s, "material", w)), as_bool(req(s, "invert", w), "Sphere.invert")};
    if (tag == "Plane")
        return PlaneData{as_vec3(req(s, "point", w), "point"), as_vec3(req(s, "normal", w), "normal"), material_from_yaml(req(s, "material", w))};
    if (tag == "Triangle")
        return TriangleData{as_vec3(req(s, "v0", w), "v0"), as_vec3(req(s, "v1", w), "v1"), as_vec3(req(s, "v2", w), "v2"),
                            material_from_yaml(req(s, "material", w))};
    if (tag == "Rectangle")
        return RectangleData{as_vec3(req(s, "corner", w), "corner"), as_vec3(req(s, "edge_a", w), "edge_a"),
                             as_vec3(req(s, "edge_b", w), "edge_b"), material_from_yaml(req(s, "material", w))};
    if (tag == "Box") return BoxData{as_vec3(req(s, "min", w), "min"), as_vec3(req(s, "max", w), "max"), material_from_yaml(req(s, "material", w))};
    if (tag == "Mesh") {
        MeshData m{{}, {}, material_from_yaml(req(s, "material", w))};
        const Node &vs = req(s, "vertices", w), &fs = req(s, "faces", w);
        if (vs.kind != Node::Seq || fs.kind != Node::Seq) throw Error("Mesh: vertices and faces must be sequences");
        for (const Node &v : vs.seq) m.vertices.push_back(as_vec3(v, "Mesh.vertices"));
        for (const Node &f : fs.seq) {
            if (f.kind != Node::Seq || f.seq.size() != 3) throw Error("Mesh.faces: expected a sequence of 3 indices");
            std::array<int64_t, 3> idx;
            for (int k = 0; k < 3; k++) {
                idx[k] = (int64_t)as_usize(f.seq[k], "Mesh.faces");
                if ((size_t)idx[k] >= m.vertices.size()) throw Error("Mesh.faces: vertex index out of range");
            }
            m.faces.push_back(idx);
        }
        return m;
    }
    throw Error("unknown variant `" + tag + "`, expected `Sphere` or `Plane`");
}

SceneData scene_from_tree(const Node &d) {
    if (d.kind != Node::Map) throw Error("SceneData: expected a map");
    const char *what = "SceneData";
    const Node &os = req(d, "output_settings", what), &cs = req(d, "camera_settings", what), &cd = req(d, "camera_data", what);
    const Node &shapes = req(d, "shapes", what);
    if (shapes.kind != Node::Seq) throw Error("shapes: expected a sequence");
    SceneData sd;
    const Node &name = req(d, "scene_name", what);
    if (name.kind != Node::Scalar) throw Error("scene_name: invalid type: expected a string");
    sd.scene_name = name.scalar;
    sd.output_settings = OutputSettings{as_usize(req(os, "image_width", "output_settings"), "output_settings.image_width"),
                                        as_usize(req(os, "image_height", "output_settings"), "output_settings.image_height"),
                                        as_f64(req(os, "pixel_size", "output_settings"), "pixel_size")};
    sd.background = as_vec3(req(d, "background", what), "background");
    for (const Node &s : shapes.seq) sd.shapes.push_back(shape_from_yaml(s));
    sd.camera_settings = CameraSettings{as_vec3(req(cs, "eye", "camera_settings"), "eye"), as_vec3(req(cs, "look_at", "camera_settings"), "look_at"),
                                        as_vec3(req(cs, "up", "camera_settings"), "up")};
    sd.camera_data = CameraData{as_f64(req(cd, "zoom_factor", "camera_data"), "zoom_factor"),
                                as_f64(req(cd, "view_plane_distance", "camera_data"), "view_plane_distance"),
                                as_f64(req(cd, "focal_distance", "camera_data"), "focal_distance"),
                                as_f64(req(cd, "lens_radius", "camera_data"), "lens_radius")};
    return sd;
}

flux_material material_to_flat(const MaterialData &m) {
    flux_material f{};
    if (auto p = std::get_if<MatteData>(&m)) {
        f.kind = FLUX_MAT_MATTE;
        std::copy(p->diffuse_color.begin(), p->diffuse_color.end(), f.color);
        f.k = p->diffuse_coefficient;
    } else if (auto p = std::get_if<EmissiveData>(&m)) {
        f.kind = FLUX_MAT_EMISSIVE;
        std::copy(p->color.begin(), p->color.end(), f.color);
        f.k = p->power;
    } else if (auto p = std::get_if<ReflectiveData>(&m)) {
        f.kind = FLUX_MAT_REFLECTIVE;
        std::copy(p->reflect_color.begin(), p->reflect_color.end(), f.color);
        f.k = p->reflect_amount;
    } else {
        const auto &g = std::get<GlossyReflectiveData>(m);
        f.kind = FLUX_MAT_GLOSSY;
        std::copy(g.reflect_color.begin(), g.reflect_color.end(), f.color);
        f.k = g.reflect_amount;
        f.exp = g.reflect_exponent;
    }
    return f;
}

Vec3 add3(const Vec3 &a, const Vec3 &b) { return Vec3{a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }

std::vector<RectangleData> box_rectangles(const BoxData &b) {
    const double x0 = b.min[0], y0 = b.min[1], z0 = b.min[2], x1 = b.max[0], y1 = b.max[1], z1 = b.max[2];
    const double dx = x1 - x0, dy = y1 - y0, dz = z1 - z0;
    const MaterialData &m = b.material;
    return {
        RectangleData{{x0, y0, z0}, {0.0, dy, 0.0}, {dx, 0.0, 0.0}, m},   // z = z0, normal -z
        RectangleData{{x0, y0, z1}, {dx, 0.0, 0.0}, {0.0, dy, 0.0}, m},   // z = z1, normal +z
        RectangleData{{x0, y0, z0}, {0.0, 0.0, dz}, {0.0, dy, 0.0}, m},   // x = x0, normal -x
        RectangleData{{x1, y0, z0}, {0.0, dy, 0.0}, {0.0, 0.0, dz}, m},   // x = x1, normal +x
        RectangleData{{x0, y0, z0}, {dx, 0.0, 0.0}, {0.0, 0.0, dz}, m},   // y = y0, normal -y
        RectangleData{{x0, y1, z0}, {0.0, 0.0, dz}, {dx, 0.0, 0.0}, m},   // y = y1, normal +y
    };
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// SceneData
// ------------------------------------------------------------------------------------------------
SceneData SceneData::from_yaml_string(const std::string &text) { return scene_from_tree(Parser(text).parse()); }

namespace detail {
SceneData scene_from_node(const Node &d) { return scene_from_tree(d); }
}  // namespace detail

SceneData SceneData::from_yaml_file(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw Error("cannot open scene file `" + path + "`");
    std::stringstream ss;
    ss << f.rdbuf();
    return from_yaml_string(ss.str());
}

SceneData SceneData::with_size(uint32_t width, uint32_t height) const {
    SceneData s = *this;
    s.output_settings.image_width = width;
    s.output_settings.image_height = height;
    return s;
}

const flux_scene_flat *FlatScene::ptr() {
    flat.n_materials = (uint32_t)materials.size();
    flat.materials = materials.data();
    flat.n_spheres = (uint32_t)sphere_radius.size();
    flat.sphere_center = sphere_center.data();
    flat.sphere_radius = sphere_radius.data();
    flat.sphere_invert = sphere_invert.data();
    flat.sphere_shape_id = sphere_shape_id.data();
    flat.sphere_material = sphere_material.data();
    flat.n_planes = (uint32_t)plane_shape_id.size();
    flat.plane_point = plane_point.data();
    flat.plane_normal = plane_normal.data();
    flat.plane_shape_id = plane_shape_id.data();
    flat.plane_material = plane_material.data();
    flat.n_triangles = (uint32_t)tri_shape_id.size();
    flat.tri_v0 = tri_v0.data();
    flat.tri_v1 = tri_v1.data();
    flat.tri_v2 = tri_v2.data();
    flat.tri_shape_id = tri_shape_id.data();
    flat.tri_material = tri_material.data();
    return &flat;
}

std::unique_ptr<FlatScene> SceneData::flatten() const {
    auto fs = std::make_unique<FlatScene>();
    FlatScene &f = *fs;
    auto mat_id = [&f](const MaterialData &m) -> uint32_t {
        const flux_material k = material_to_flat(m);
        for (size_t i = 0; i < f.materials.size(); i++) {
            const flux_material &q = f.materials[i];
            if (q.kind == k.kind && q.color[0] == k.color[0] && q.color[1] == k.color[1] && q.color[2] == k.color[2] && q.k == k.k && q.exp == k.exp)
                return (uint32_t)i;
        }
        f.materials.push_back(k);
        return (uint32_t)f.materials.size() - 1;
    };
    auto push3 = [](std::vector<double> &v, const Vec3 &a) { v.insert(v.end(), a.begin(), a.end()); };
    uint32_t shape_id = 0;
    auto push_tri = [&](const Vec3 &a, const Vec3 &b, const Vec3 &c, uint32_t mat) {
        push3(f.tri_v0, a);
        push3(f.tri_v1, b);
        push3(f.tri_v2, c);
        f.tri_shape_id.push_back(shape_id++);
        f.tri_material.push_back(mat);
    };
    auto push_rect = [&](const RectangleData &r) {
        const uint32_t mat = mat_id(r.material);
        const Vec3 pa = add3(r.corner, r.edge_a), pab = add3(pa, r.edge_b), pb = add3(r.corner, r.edge_b);
        push_tri(r.corner, pa, pab, mat);   // both wound so that the normal is edge_a x edge_b
        push_tri(r.corner, pab, pb, mat);
    };
    for (const ShapeData &sh : shapes) {
        if (auto s = std::get_if<SphereData>(&sh)) {
            push3(f.sphere_center, s->center);
            f.sphere_radius.push_back(s->radius);
            f.sphere_invert.push_back(s->invert ? 1 : 0);
            f.sphere_shape_id.push_back(shape_id++);
            f.sphere_material.push_back(mat_id(s->material));
        } else if (auto p = std::get_if<PlaneData>(&sh)) {
            push3(f.plane_point, p->point);
            push3(f.plane_normal, p->normal);
            f.plane_shape_id.push_back(shape_id++);
            f.plane_material.push_back(mat_id(p->material));
        } else if (auto t = std::get_if<TriangleData>(&sh)) {
            push_tri(t->v0, t->v1, t->v2, mat_id(t->material));
        } else if (auto r = std::get_if<RectangleData>(&sh)) {
            push_rect(*r);
        } else if (auto b = std::get_if<BoxData>(&sh)) {
            for (const RectangleData &r : box_rectangles(*b)) push_rect(r);
        } else {
            const MeshData &m = std::get<MeshData>(sh);
            const uint32_t mat = mat_id(m.material);
            const size_t nf = m.faces.size(), nv = m.vertices.size();
            for (std::vector<double> *v : {&f.tri_v0, &f.tri_v1, &f.tri_v2}) v->reserve(v->size() + 3 * nf);
            f.tri_shape_id.reserve(f.tri_shape_id.size() + nf);
            f.tri_material.reserve(f.tri_material.size() + nf);
            for (const auto &fc : m.faces) {
                // the YAML reader has checked this; a MeshData built in code has not
                for (int k = 0; k < 3; k++)
                    if (fc[k] < 0 || (size_t)fc[k] >= nv) throw Error("Mesh.faces: vertex index out of range");
                push_tri(m.vertices[fc[0]], m.vertices[fc[1]], m.vertices[fc[2]], mat);
            }
        }
    }
    f.n_shapes = shape_id;
    flux_scene_flat &s = f.flat;
    s.image_width = output_settings.image_width;
    s.image_height = output_settings.image_height;
    s.pixel_size = output_settings.pixel_size;
    for (int k = 0; k < 3; k++) {
        s.background[k] = background[k];
        s.eye[k] = camera_settings.eye[k];
        s.look_at[k] = camera_settings.look_at[k];
        s.up[k] = camera_settings.up[k];
    }
    s.zoom_factor = camera_data.zoom_factor;
    s.view_plane_distance = camera_data.view_plane_distance;
    s.focal_distance = camera_data.focal_distance;
    s.lens_radius = camera_data.lens_radius;
    f.ptr();
    return fs;
}

// Flattened scene as text, doubles in hex-float so the comparison with the Python loader is exact.
void dump_flat_text(const std::string &path, FlatScene &f) {
    FILE *o = std::fopen(path.c_str(), "w");
    if (!o) throw Error("cannot write `" + path + "`");
    const flux_scene_flat &s = *f.ptr();
    auto dv = [&](const char *name, const double *p, size_t n) {
        std::fprintf(o, "%s %zu", name, n);
        for (size_t i = 0; i < n; i++) std::fprintf(o, " %a", p[i]);
        std::fprintf(o, "\n");
    };
    auto uv = [&](const char *name, const uint32_t *p, size_t n) {
        std::fprintf(o, "%s %zu", name, n);
        for (size_t i = 0; i < n; i++) std::fprintf(o, " %u", p[i]);
        std::fprintf(o, "\n");
    };
    std::fprintf(o, "image %u %u\n", s.image_width, s.image_height);
    const double scal[] = {s.pixel_size, s.zoom_factor, s.view_plane_distance, s.focal_distance, s.lens_radius};
    dv("scalars", scal, 5);
    dv("background", s.background, 3);
    dv("eye", s.eye, 3);
    dv("look_at", s.look_at, 3);
    dv("up", s.up, 3);
    std::fprintf(o, "materials %u\n", s.n_materials);
    for (uint32_t i = 0; i < s.n_materials; i++) {
        const flux_material &m = s.materials[i];
        std::fprintf(o, "material %u %a %a %a %a %a\n", m.kind, m.color[0], m.color[1], m.color[2], m.k, m.exp);
    }
    dv("sphere_center", s.sphere_center, 3 * (size_t)s.n_spheres);
    dv("sphere_radius", s.sphere_radius, s.n_spheres);
    std::fprintf(o, "sphere_invert %u", s.n_spheres);
    for (uint32_t i = 0; i < s.n_spheres; i++) std::fprintf(o, " %u", (unsigned)s.sphere_invert[i]);
    std::fprintf(o, "\n");
    uv("sphere_shape_id", s.sphere_shape_id, s.n_spheres);
    uv("sphere_material", s.sphere_material, s.n_spheres);
    dv("plane_point", s.plane_point, 3 * (size_t)s.n_planes);
    dv("plane_normal", s.plane_normal, 3 * (size_t)s.n_planes);
    uv("plane_shape_id", s.plane_shape_id, s.n_planes);
    uv("plane_material", s.plane_material, s.n_planes);
    dv("tri_v0", s.tri_v0, 3 * (size_t)s.n_triangles);
    dv("tri_v1", s.tri_v1, 3 * (size_t)s.n_triangles);
    dv("tri_v2", s.tri_v2, 3 * (size_t)s.n_triangles);
    uv("tri_shape_id", s.tri_shape_id, s.n_triangles);
    uv("tri_material", s.tri_material, s.n_triangles);
    std::fclose(o);
}


std::vector<WorkUnit> work_units(uint32_t image_height, uint32_t rows_per_work_unit, uint64_t job_id) {
    if (rows_per_work_unit == 0) throw Error("rows_per_work_unit must be >= 1");
    std::vector<WorkUnit> u;
    for (uint32_t r = 0; r < image_height; r += rows_per_work_unit)
        u.push_back(WorkUnit{r, std::min(image_height, r + rows_per_work_unit) - 1, job_id});
    return u;
}

// ------------------------------------------------------------------------------------------------
// Image
// ------------------------------------------------------------------------------------------------
void Image::set_rows(const WorkUnitResult &r) {
    const size_t row_elems = (size_t)width * 3;
    const size_t n = (size_t)(r.work_unit.row_end - r.work_unit.row_start + 1) * row_elems;
    if (r.work_unit.row_end >= height || r.rows.size() != n) throw Error("Image::set_rows: result does not fit the image");
    std::copy(r.rows.begin(), r.rows.end(), pixels.begin() + (size_t)r.work_unit.row_start * row_elems);
}

void Image::write(const std::string &path) const {
    if (flux_write_ppm(path.c_str(), width, height, pixels.data()) != FLUX_OK) throw Error("cannot write `" + path + "`");
}

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
GpuContext::GpuContext(int device) : device_(device) {
    const int rc = flux_ctx_create(device, &ctx_);
    if (rc != FLUX_OK) throw Error(std::string("flux_ctx_create: ") + flux_last_error(nullptr));
}
GpuContext::~GpuContext() {
    if (ctx_) flux_ctx_destroy(ctx_);
}
void GpuContext::check(int rc, const char *what) const {
    if (rc != FLUX_OK) throw Error(std::string(what) + ": " + flux_last_error(ctx_));
}

Scene Scene::from_data(const SceneData &sd, const JobConfiguration &cfg) {
    Scene s;
    s.data = sd;
    s.job_config = cfg;
    s.flat = sd.flatten();
    return s;
}

Camera Camera::create(GpuContext &ctx, const Scene &scene, const JobConfiguration &cfg, uint32_t num_sets, uint64_t seed) {
    flux_job_config jc{cfg.sample_root, cfg.max_trace_depth, cfg.rows_per_work_unit};
    ctx.check(flux_set_scene(ctx.get(), scene.flat->ptr(), &jc), "flux_set_scene");
    ctx.check(flux_generate_samples(ctx.get(), seed, num_sets), "flux_generate_samples");
    return Camera(ctx, scene.data.output_settings.image_width, scene.data.output_settings.image_height);
}

WorkUnitResult Camera::render(const Scene &, const WorkUnit &unit) const {
    if (unit.row_end < unit.row_start) throw Error("Camera::render: row_end < row_start");
    WorkUnitResult r{unit, std::vector<double>((size_t)(unit.row_end - unit.row_start + 1) * width_ * 3)};
    ctx_->check(flux_render_rows(ctx_->get(), unit.row_start, unit.row_end, r.rows.data()), "flux_render_rows");
    return r;
}

std::vector<double> Camera::render_row_list(const std::vector<uint32_t> &rows) const {
    std::vector<double> out(rows.size() * (size_t)width_ * 3);
    ctx_->check(flux_render_row_list(ctx_->get(), rows.data(), (uint32_t)rows.size(), out.data()), "flux_render_row_list");
    return out;
}

WorkUnitResult Camera::render_progressive(const WorkUnit &unit, uint32_t sample_root, uint32_t batch,
                                          const std::function<bool(uint32_t, const WorkUnitResult &)> &on_pass) const {
    if (unit.row_end < unit.row_start) throw Error("Camera::render_progressive: row_end < row_start");
    if (batch == 0) throw Error("Camera::render_progressive: batch must be >= 1");
    std::vector<uint32_t> rows;
    for (uint32_t r = unit.row_start; r <= unit.row_end; r++) rows.push_back(r);
    ctx_->check(flux_progressive_begin(ctx_->get(), rows.data(), (uint32_t)rows.size()), "flux_progressive_begin");
    WorkUnitResult r{unit, std::vector<double>(rows.size() * (size_t)width_ * 3)};
    const uint32_t n = sample_root * sample_root;
    for (uint32_t done = 0; done < n;) {
        const uint32_t end = std::min(n, done + batch);
        ctx_->check(flux_progressive_pass(ctx_->get(), done, end, r.rows.data()), "flux_progressive_pass");
        done = end;
        if (on_pass && !on_pass(done, r)) break;
    }
    return r;
}

float Camera::last_kernel_ms() const {
    float ms = 0.f;
    flux_last_kernel_ms(ctx_->get(), &ms);
    return ms;
}

GpuWorker::GpuWorker(std::vector<int> devices, uint64_t seed, uint32_t tile_rows)
    : devices_(std::move(devices)), seed_(seed), tile_rows_(tile_rows) {
    if (devices_.empty()) throw Error("GpuWorker: no devices");
    if (tile_rows_ == 0) throw Error("GpuWorker: tile_rows must be >= 1");
    // device contexts belong to the worker, not to a job (workers.rs:26-41 builds the thread pool once)
    for (int d : devices_) contexts_.push_back(std::make_unique<GpuContext>(d));
}

WorkerInfo GpuWorker::info() const {
    std::string name = "gpu";
    for (size_t i = 0; i < devices_.size(); i++) name += (i ? "," : "") + std::to_string(devices_[i]);
    return WorkerInfo{name, (uint32_t)devices_.size()};
}

std::vector<WorkUnitResult> GpuWorker::run_job(const SceneData &sd, const JobConfiguration &cfg, const std::atomic<bool> *cancel) {
    Scene scene = Scene::from_data(sd, cfg);
    GpuContext &ctx = *contexts_[0];
    Camera camera = Camera::create(ctx, scene, cfg, sd.output_settings.image_width, seed_);
    std::vector<WorkUnitResult> out;
    for (const WorkUnit &u : work_units(sd.output_settings.image_height, cfg.rows_per_work_unit)) {
        if (cancel && cancel->load()) break;   // stop issuing; nothing is in flight between units
        out.push_back(camera.render(scene, u));
    }
    return out;
}

Image GpuWorker::render_job_progressive(const SceneData &sd, const JobConfiguration &cfg, uint32_t batch,
                                        const std::function<bool(uint32_t, const Image &)> &on_pass) {
    if (batch == 0) throw Error("render_job_progressive: batch must be >= 1");
    const uint32_t W = sd.output_settings.image_width, H = sd.output_settings.image_height;
    const uint32_t n = cfg.sample_root * cfg.sample_root;
    const uint32_t world = (uint32_t)devices_.size();
    const size_t row_elems = (size_t)W * 3;
    Image img(W, H);
    Scene scene = Scene::from_data(sd, cfg);
    std::vector<std::vector<uint32_t>> rows(world);
    std::vector<std::vector<double>> px(world);
    std::vector<std::string> errors(world);
    auto on_every_gpu = [&](const std::function<void(uint32_t)> &f) {
        auto guarded = [&](uint32_t rank) {
            try {
                f(rank);
            } catch (const std::exception &e) {
                errors[rank] = e.what();
            }
        };
        if (world == 1) {
            guarded(0);
        } else {
            std::vector<std::thread> th;
            for (uint32_t r = 0; r < world; r++) th.emplace_back(guarded, r);
            for (auto &t : th) t.join();
        }
        for (const std::string &e : errors)
            if (!e.empty()) throw Error(e);
    };
    on_every_gpu([&](uint32_t rank) {   // Scene::from_data + Camera::new, then name this GPU's rows
        uint32_t cnt = 0;
        flux_shard_rows(H, tile_rows_, rank, world, nullptr, &cnt);
        rows[rank].resize(cnt);
        flux_shard_rows(H, tile_rows_, rank, world, rows[rank].data(), &cnt);
        px[rank].resize(cnt * row_elems);
        GpuContext &ctx = *contexts_[rank];
        Camera::create(ctx, scene, cfg, W, seed_);
        ctx.check(flux_progressive_begin(ctx.get(), rows[rank].data(), cnt), "flux_progressive_begin");
    });
    for (uint32_t done = 0; done < n;) {
        const uint32_t end = std::min(n, done + batch);
        on_every_gpu([&](uint32_t rank) {
            if (rows[rank].empty()) return;
            GpuContext &ctx = *contexts_[rank];
            ctx.check(flux_progressive_pass(ctx.get(), done, end, px[rank].data()), "flux_progressive_pass");
            for (size_t k = 0; k < rows[rank].size(); k++)   // disjoint rows: no synchronisation needed
                std::copy(px[rank].begin() + k * row_elems, px[rank].begin() + (k + 1) * row_elems,
                          img.pixels.begin() + (size_t)rows[rank][k] * row_elems);
        });
        done = end;
        if (on_pass && !on_pass(done, img)) break;
    }
    return img;
}

void GpuWorker::begin_job(const SceneData &sd, const JobConfiguration &cfg) {
    auto job = std::make_unique<ActiveJob>();
    job->scene = Scene::from_data(sd, cfg);
    job->width = sd.output_settings.image_width;
    job->height = sd.output_settings.image_height;
    const uint32_t world = (uint32_t)devices_.size();
    std::vector<std::string> errors(world);
    std::vector<std::unique_ptr<Camera>> cams(world);
    auto setup = [&](uint32_t rank) {
        try {
            cams[rank] = std::make_unique<Camera>(Camera::create(*contexts_[rank], job->scene, cfg, job->width, seed_));
        } catch (const std::exception &e) {
            errors[rank] = e.what();
        }
    };
    if (world == 1) {
        setup(0);
    } else {
        std::vector<std::thread> th;
        for (uint32_t r = 0; r < world; r++) th.emplace_back(setup, r);
        for (auto &t : th) t.join();
    }
    for (const std::string &e : errors)
        if (!e.empty()) throw Error(e);
    for (auto &c : cams) job->cameras.push_back(*c);
    job_ = std::move(job);
}

WorkUnitResult GpuWorker::render_unit(const WorkUnit &unit) {
    if (!job_) throw Error("GpuWorker::render_unit: no job");
    if (unit.row_end < unit.row_start || unit.row_end >= job_->height) throw Error("GpuWorker::render_unit: rows outside the image");
    const uint32_t world = (uint32_t)devices_.size();
    if (world == 1) return job_->cameras[0].render(job_->scene, unit);
    const size_t row_elems = (size_t)job_->width * 3;
    WorkUnitResult res{unit, std::vector<double>((size_t)(unit.row_end - unit.row_start + 1) * row_elems)};
    std::vector<std::string> errors(world);
    auto shard = [&](uint32_t rank) {
        try {
            std::vector<uint32_t> rows;   // same ownership rule as flux_shard_rows, restricted to the unit
            for (uint32_t r = unit.row_start; r <= unit.row_end; r++)
                if ((r / tile_rows_) % world == rank) rows.push_back(r);
            if (rows.empty()) return;
            const std::vector<double> px = job_->cameras[rank].render_row_list(rows);
            for (size_t k = 0; k < rows.size(); k++)
                std::copy(px.begin() + k * row_elems, px.begin() + (k + 1) * row_elems,
                          res.rows.begin() + (size_t)(rows[k] - unit.row_start) * row_elems);
        } catch (const std::exception &e) {
            errors[rank] = e.what();
        }
    };
    std::vector<std::thread> th;
    for (uint32_t r = 0; r < world; r++) th.emplace_back(shard, r);
    for (auto &t : th) t.join();
    for (const std::string &e : errors)
        if (!e.empty()) throw Error(e);
    return res;
}

Image GpuWorker::render_job(const SceneData &sd, const JobConfiguration &cfg, double *render_seconds) {
    const uint32_t W = sd.output_settings.image_width, H = sd.output_settings.image_height;
    Image img(W, H);
    Scene scene = Scene::from_data(sd, cfg);
    const uint32_t world = (uint32_t)devices_.size();
    std::vector<std::string> errors(world);
    // the reference's timer covers Scene::from_data + Camera::new + all rendering (manager.rs:145-170)
    const auto t0 = std::chrono::steady_clock::now();
    // ONE frame on the first GPU; every GPU's render kernel stores its rows into it through NVLink peer memory
    // (flux_frame_*: replaces the RowsReady stream into the manager's ImageBuilder, manager.rs:100,156-162) — no
    // per-GPU slices in host memory, no assembly on the host; one device-to-host copy at the end.
    GpuContext &owner = *contexts_[0];
    flux_frame *frame = nullptr;
    owner.check(flux_frame_create(owner.get(), W, H, &frame), "flux_frame_create");
    std::vector<flux_frame *> views(world, nullptr);
    auto shard = [&](uint32_t rank) {
        try {
            uint32_t n = 0;
            flux_shard_rows(H, tile_rows_, rank, world, nullptr, &n);
            std::vector<uint32_t> rows(n);
            flux_shard_rows(H, tile_rows_, rank, world, rows.data(), &n);
            if (n == 0) return;
            GpuContext &ctx = *contexts_[rank];
            Camera::create(ctx, scene, cfg, W, seed_);   // same seed on every GPU: identical sample sets
            ctx.check(flux_frame_open_peer(ctx.get(), frame, &views[rank]), "flux_frame_open_peer");
            ctx.check(flux_render_row_list_into_frame(ctx.get(), rows.data(), n, views[rank], nullptr), "flux_render_row_list_into_frame");
            ctx.check(flux_ctx_sync(ctx.get()), "flux_ctx_sync");
        } catch (const std::exception &e) {
            errors[rank] = e.what();
        }
    };
    if (world == 1) {
        shard(0);
    } else {
        std::vector<std::thread> th;
        for (uint32_t r = 0; r < world; r++) th.emplace_back(shard, r);
        for (auto &t : th) t.join();   // the join is the barrier: every GPU's rows are in the frame
    }
    std::string first_error;
    for (const std::string &e : errors)
        if (!e.empty() && first_error.empty()) first_error = e;
    int rc = FLUX_OK;
    if (first_error.empty()) rc = flux_frame_read(frame, img.pixels.data());
    for (flux_frame *v : views)
        if (v) flux_frame_close(v);
    flux_frame_close(frame);
    if (!first_error.empty()) throw Error(first_error);
    if (rc != FLUX_OK) throw Error(std::string("flux_frame_read: ") + flux_last_error(nullptr));
    if (render_seconds) *render_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return img;
}

}  // namespace flux
