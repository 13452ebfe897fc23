// cbor.cpp — see cbor.hpp.
#include "cbor.hpp"

#include <cmath>
#include <cstring>

namespace flux {
namespace cbor {

using detail::Node;

namespace {
constexpr int MAX_DEPTH = 64;
// Every item becomes a detail::Node (~100 bytes with its string, vector and map), so the item budget is a memory
// budget: 2^26 items bound one message at a few GB where 2^28 let a 256 MB message of one-byte items ask for tens of GB
// (ADVICE r1).  The largest legitimate messages: SetJob with a million-triangle mesh, ~1.5e7 items; RowsReady of a
// 50-row unit at 16 K pixels per row, 3.3e6.
constexpr uint64_t MAX_ITEMS = 1ull << 26;
constexpr uint64_t MAX_STRING = 1ull << 28;

double half_to_double(uint16_t h) {
    const int e = (h >> 10) & 0x1f, m = h & 0x3ff;
    double v;
    if (e == 0) v = std::ldexp((double)m, -24);
    else if (e != 31) v = std::ldexp((double)(m + 1024), e - 25);
    else v = m == 0 ? INFINITY : NAN;
    return (h & 0x8000) ? -v : v;
}
}  // namespace

size_t MemorySource::read(uint8_t *dst, size_t n) {
    const size_t k = n < left ? n : left;
    std::memcpy(dst, p, k);
    p += k;
    left -= k;
    return k;
}

bool Reader::fill() {
    pos_ = 0;
    end_ = src_.read(buf_, sizeof buf_);
    return end_ != 0;
}

bool Reader::byte(uint8_t &b) {
    if (pos_ == end_ && !fill()) return false;
    b = buf_[pos_++];
    return true;
}

uint8_t Reader::need() {
    uint8_t b;
    if (!byte(b)) throw Error("cbor: unexpected end of stream inside an item");
    return b;
}

void Reader::need(uint8_t *dst, size_t n) {
    while (n) {
        if (pos_ == end_ && !fill()) throw Error("cbor: unexpected end of stream inside an item");
        const size_t k = std::min(n, end_ - pos_);
        std::memcpy(dst, buf_ + pos_, k);
        pos_ += k;
        dst += k;
        n -= k;
    }
}

uint64_t Reader::argument(uint8_t info) {
    if (info < 24) return info;
    if (info > 27) throw Error("cbor: reserved additional information " + std::to_string(info));
    const int nbytes = 1 << (info - 24);
    uint8_t b[8];
    need(b, nbytes);
    uint64_t v = 0;
    for (int i = 0; i < nbytes; i++) v = (v << 8) | b[i];
    return v;
}

std::string Reader::bytes(uint8_t major, uint8_t info) {
    std::string s;
    auto chunk = [&](uint64_t n) {
        if (n > MAX_STRING || s.size() + n > MAX_STRING) throw Error("cbor: string too long");
        const size_t at = s.size();
        s.resize(at + (size_t)n);
        need(reinterpret_cast<uint8_t *>(&s[at]), (size_t)n);
    };
    if (info != 31) {
        chunk(argument(info));
        return s;
    }
    for (;;) {   // indefinite length: definite chunks of the same major type until the break
        const uint8_t c = need();
        if (c == 0xff) return s;
        if ((c >> 5) != major || (c & 31) == 31) throw Error("cbor: bad chunk inside an indefinite-length string");
        chunk(argument(c & 31));
    }
}

bool Reader::next(Node &out) {
    uint8_t first;
    if (!byte(first)) return false;
    budget_ = MAX_ITEMS;
    out = Node();
    item(first, out, 0);
    return true;
}

void Reader::item(uint8_t first, Node &out, int depth) {
    if (depth > MAX_DEPTH) throw Error("cbor: nesting too deep");
    if (budget_-- == 0) throw Error("cbor: document too large");
    const uint8_t major = first >> 5, info = first & 31;
    out.cbor = true;
    switch (major) {
    case 0:
        out.kind = Node::Scalar;
        out.bin = Node::U64;
        out.u = argument(info);
        return;
    case 1: {   // -1 - n
        const uint64_t n = argument(info);
        if (n > 0x7fffffffffffffffull) throw Error("cbor: negative integer out of range");
        out.kind = Node::Scalar;
        out.bin = Node::I64;
        out.u = (uint64_t)(-1 - (int64_t)n);
        return;
    }
    case 2:
    case 3:
        out.kind = Node::Scalar;
        out.bin = Node::Text;
        out.quoted = true;
        out.scalar = bytes(major, info);
        return;
    case 4: {
        out.kind = Node::Seq;
        if (info == 31) {
            for (;;) {
                const uint8_t c = need();
                if (c == 0xff) return;
                out.seq.emplace_back();
                item(c, out.seq.back(), depth + 1);
            }
        }
        const uint64_t n = argument(info);
        if (n > budget_) throw Error("cbor: array longer than the document may be");
        out.seq.reserve((size_t)std::min<uint64_t>(n, 1u << 16));
        for (uint64_t i = 0; i < n; i++) {
            out.seq.emplace_back();
            item(need(), out.seq.back(), depth + 1);
        }
        return;
    }
    case 5: {
        out.kind = Node::Map;
        const bool indefinite = info == 31;
        const uint64_t n = indefinite ? ~0ull : argument(info);
        if (!indefinite && n > budget_) throw Error("cbor: map longer than the document may be");
        for (uint64_t i = 0; i < n; i++) {
            const uint8_t c = need();
            if (indefinite && c == 0xff) return;
            Node key;
            item(c, key, depth + 1);
            // serde_cbor's default serializer writes struct fields and enum variants as text keys
            if (key.kind != Node::Scalar || key.bin != Node::Text) throw Error("cbor: map key is not a text string");
            out.map.emplace_back(std::move(key.scalar), Node());
            item(need(), out.map.back().second, depth + 1);
        }
        return;
    }
    case 6:   // tag: the tagged item stands for itself
        argument(info);
        item(need(), out, depth + 1);
        return;
    default:   // 7
        out.kind = Node::Scalar;
        if (info == 20 || info == 21) {
            out.bin = Node::Bool;
            out.u = info == 21;
        } else if (info == 22 || info == 23) {
            out.kind = Node::Null;
        } else if (info == 25) {
            out.bin = Node::F64;
            out.f = half_to_double((uint16_t)argument(info));
        } else if (info == 26) {
            const uint32_t bits = (uint32_t)argument(info);
            float f;
            std::memcpy(&f, &bits, 4);
            out.bin = Node::F64;
            out.f = (double)f;
        } else if (info == 27) {
            const uint64_t bits = argument(info);
            out.bin = Node::F64;
            std::memcpy(&out.f, &bits, 8);
        } else if (info == 31) {
            throw Error("cbor: break outside an indefinite-length item");
        } else {
            throw Error("cbor: unsupported simple value");
        }
        return;
    }
}

void Writer::head(uint8_t major, uint64_t v) {
    const uint8_t m = (uint8_t)(major << 5);
    int nbytes;
    if (v < 24) {
        out.push_back((char)(m | v));
        return;
    } else if (v <= 0xff) {
        out.push_back((char)(m | 24));
        nbytes = 1;
    } else if (v <= 0xffff) {
        out.push_back((char)(m | 25));
        nbytes = 2;
    } else if (v <= 0xffffffffull) {
        out.push_back((char)(m | 26));
        nbytes = 4;
    } else {
        out.push_back((char)(m | 27));
        nbytes = 8;
    }
    for (int i = nbytes - 1; i >= 0; i--) out.push_back((char)(v >> (8 * i)));
}

void Writer::f64(double v) {
    if (std::isnan(v)) {
        out.append("\xf9\x7e\x00", 3);
    } else if (std::isinf(v)) {
        out.append(v > 0 ? "\xf9\x7c\x00" : "\xf9\xfc\x00", 3);
    } else if ((double)(float)v == v) {   // exact in f32 (includes +-0 and f32 subnormals)
        const float f = (float)v;
        uint32_t bits;
        std::memcpy(&bits, &f, 4);
        out.push_back((char)0xfa);
        for (int i = 3; i >= 0; i--) out.push_back((char)(bits >> (8 * i)));
    } else {
        uint64_t bits;
        std::memcpy(&bits, &v, 8);
        out.push_back((char)0xfb);
        for (int i = 7; i >= 0; i--) out.push_back((char)(bits >> (8 * i)));
    }
}

}  // namespace cbor
}  // namespace flux
