// node.hpp — the document tree both readers produce: the YAML reader (scene files, flux/src/main.rs:28-29) and the
// CBOR reader (network jobs, fluxcore/src/workers.rs:106-110).  serde drives both formats through the same
// Deserialize impls in the reference; here both build a Node and the same scene_from_node() consumes it.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace flux {
namespace detail {

struct Node {
    enum Kind { Null, Scalar, Seq, Map } kind = Null;
    // a YAML scalar is text that the accessor parses; a CBOR scalar arrives typed and keeps its exact value
    enum Bin { Text, F64, U64, I64, Bool } bin = Text;
    std::string scalar;
    bool quoted = false;   // text that must not be read as a number / boolean
    bool cbor = false;     // built by the CBOR reader: enums may come in serde_cbor's array form [variant, content]
    double f = 0.0;
    uint64_t u = 0;        // U64 value, I64 two's complement, Bool 0/1
    std::vector<Node> seq;
    std::vector<std::pair<std::string, Node>> map;

    const Node *find(const std::string &key) const {
        if (kind != Map) return nullptr;
        for (auto &kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

}  // namespace detail

struct SceneData;
namespace detail {
// serde's Deserialize for SceneData over either format: every field required, unknown keys ignored,
// enums externally tagged, Vector3 / Point3 as 3-sequences, Color as a 3-sequence or {r, g, b}
SceneData scene_from_node(const Node &d);
}  // namespace detail
}  // namespace flux
