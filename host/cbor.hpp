// cbor.hpp — the subset of RFC 7049 that serde_cbor 0.9 (Cargo.lock: serde_cbor 0.9.0) puts on the wire for the
// reference's network protocol (fluxcore/src/workers.rs:106-110, flux-node/src/main.rs:21-94): self-delimiting
// items written back to back on a TCP stream (serde_cbor::to_writer / StreamDeserializer).
//
// Reader: any well-formed item (all major types, definite and indefinite lengths, f16/f32/f64, tags skipped) into
// the same detail::Node tree the YAML reader builds, so that one scene_from_node() serves both formats.
// Writer: the forms serde_cbor's default (non-packed) serializer chooses — structs as maps with text keys in
// declaration order, sequences and tuples as definite-length arrays, unsigned integers in the shortest form,
// f64 as f32 when that is exact (f16 for NaN and the infinities), otherwise f64.  Enums: a unit variant is its
// name as text; a newtype variant is the 2-array [name, content] in serde_cbor < 0.10 (the reference pins 0.9.0;
// later releases call it "legacy") and the one-entry map {name: content} from 0.10 on — the writer does either
// (Writer::legacy_enums), the reader takes both.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "fluxhost.hpp"
#include "node.hpp"

namespace flux {
namespace cbor {

struct ByteSource {
    virtual ~ByteSource() = default;
    // up to n bytes into dst; 0 = end of stream; throws flux::Error on an I/O error
    virtual size_t read(uint8_t *dst, size_t n) = 0;
};

struct MemorySource : ByteSource {
    const uint8_t *p;
    size_t left;
    MemorySource(const void *data, size_t n) : p(static_cast<const uint8_t *>(data)), left(n) {}
    size_t read(uint8_t *dst, size_t n) override;
};

class Reader {
  public:
    explicit Reader(ByteSource &src) : src_(src) {}
    // the next item of the stream; false at a clean end of stream (no byte of a new item read);
    // throws flux::Error on a truncated or malformed item
    bool next(detail::Node &out);

  private:
    bool fill();
    bool byte(uint8_t &b);
    uint8_t need();
    void need(uint8_t *dst, size_t n);
    uint64_t argument(uint8_t info);
    void item(uint8_t first, detail::Node &out, int depth);
    std::string bytes(uint8_t major, uint8_t info);
    ByteSource &src_;
    uint8_t buf_[1 << 16];
    size_t pos_ = 0, end_ = 0;
    uint64_t budget_ = 0;   // items left before the document is rejected as unreasonably large
};

class Writer {
  public:
    std::string out;
    bool legacy_enums = true;   // [name, content] (serde_cbor < 0.10) instead of {name: content}
    // the head of a newtype variant; its content follows
    void variant(const char *name) {
        if (legacy_enums) array(2);
        else map(1);
        text(name);
    }
    void uint(uint64_t v) { head(0, v); }
    void array(uint64_t n) { head(4, n); }
    void map(uint64_t n) { head(5, n); }
    void text(const std::string &s) {
        head(3, s.size());
        out += s;
    }
    void boolean(bool b) { out.push_back(b ? (char)0xf5 : (char)0xf4); }
    void f64(double v);
    // "key": the text key of a struct field
    Writer &key(const char *k) {
        text(k);
        return *this;
    }

  private:
    void head(uint8_t major, uint64_t v);
};

}  // namespace cbor
}  // namespace flux
