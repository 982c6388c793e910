// wire.hpp -- the protobuf wire format of the per-family Shared / Group messages, hand-rolled (header only).
//
// The reference's hyperparam_bag_t / suffstats_bag_t are serialized protobuf messages
// (models/distributions.hpp:300-314 get_ss / set_ss, :355-369 get_hp / set_hp -> protobuf_to_string / ParseFromString of
// distributions' Shared / Group messages; in-tree models: microscopes/io/schema.proto:6-28).  gpu_group / gpu_hypers
// emit and parse exactly those bytes, so a bag written here can be read by the reference, by this repo's Python host
// (common_b200/wire.py, whose bytes tests/test_wire.py holds equal to the protobuf runtime's) and back.
// Field numbers: the same table as common_b200/wire.py (the un-vendored distributions schema is restated, unpinned).
// proto2 rules used: float = fixed32, uint32 = varint, repeated fields unpacked on output, packed or unpacked on input,
// unknown fields skipped.
#ifndef MICROSCOPES_B200_WIRE_HPP
#define MICROSCOPES_B200_WIRE_HPP

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../mscope_b200.h"

namespace microscopes {
namespace models {
namespace b200 {
namespace wire {

enum kind_t { FLOAT, UINT };
struct field_spec {
  int number;      // protobuf field number
  kind_t kind;
  bool repeated;
  size_t off;      // offset into the flat field vector (the order listed at enum msb_family)
  size_t cnt;      // elements (1 unless repeated)
};

// Group message <-> flat suffstat vector (reference representation: nich count / mean / count_times_variance)
inline std::vector<field_spec> group_fields(const msb_model_desc &m) {
  const size_t d = m.dim;
  switch (m.family) {
    case MSB_FAMILY_BB: return {{1, UINT, false, 0, 1}, {2, UINT, false, 1, 1}};
    case MSB_FAMILY_BNB: return {{1, UINT, false, 0, 1}, {2, UINT, false, 1, 1}};
    case MSB_FAMILY_GP: return {{1, UINT, false, 0, 1}, {2, UINT, false, 1, 1}, {3, FLOAT, false, 2, 1}};
    case MSB_FAMILY_NICH: return {{1, UINT, false, 0, 1}, {2, FLOAT, false, 1, 1}, {3, FLOAT, false, 2, 1}};
    case MSB_FAMILY_DD: return {{1, UINT, true, 1, d}};                                   // counts; count_sum is derived
    case MSB_FAMILY_NIW: return {{1, UINT, false, 0, 1}, {2, FLOAT, true, 1, d}, {3, FLOAT, true, 1 + d, d * d}};
    case MSB_FAMILY_BBNC: return {{1, FLOAT, false, 0, 1}, {2, UINT, false, 1, 1}, {3, UINT, false, 2, 1}};  // schema.proto:14-18
    case MSB_FAMILY_DM: return {{1, UINT, true, 0, d}, {2, FLOAT, false, d, 1}};          // schema.proto:25-28
  }
  throw std::runtime_error("unknown family");
}
inline std::vector<field_spec> shared_fields(const msb_model_desc &m) {
  const size_t d = m.dim;
  switch (m.family) {
    case MSB_FAMILY_BB: case MSB_FAMILY_BBNC: return {{1, FLOAT, false, 0, 1}, {2, FLOAT, false, 1, 1}};
    case MSB_FAMILY_BNB: return {{1, FLOAT, false, 0, 1}, {2, FLOAT, false, 1, 1}, {3, UINT, false, 2, 1}};
    case MSB_FAMILY_GP: return {{1, FLOAT, false, 0, 1}, {2, FLOAT, false, 1, 1}};
    case MSB_FAMILY_NICH: return {{1, FLOAT, false, 0, 1}, {2, FLOAT, false, 1, 1}, {3, FLOAT, false, 2, 1}, {4, FLOAT, false, 3, 1}};
    case MSB_FAMILY_DD: case MSB_FAMILY_DM: return {{1, FLOAT, true, 0, d}};
    case MSB_FAMILY_NIW: return {{1, FLOAT, true, 0, d}, {2, FLOAT, false, d, 1}, {3, FLOAT, true, d + 1, d * d}, {4, FLOAT, false, d + 1 + d * d, 1}};
  }
  throw std::runtime_error("unknown family");
}

inline void put_varint(std::string &out, uint64_t v) {
  while (v >= 0x80) { out.push_back((char)((v & 0x7F) | 0x80)); v >>= 7; }
  out.push_back((char)v);
}
inline uint64_t get_varint(const std::string &buf, size_t &pos) {
  uint64_t v = 0;
  for (int shift = 0; shift <= 63; shift += 7) {
    if (pos >= buf.size()) throw std::runtime_error("truncated message");
    const uint8_t b = (uint8_t)buf[pos++];
    v |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) return v;
  }
  throw std::runtime_error("malformed varint");
}

inline std::string encode(const std::vector<field_spec> &spec, const std::vector<double> &flat) {
  std::string out;
  for (const auto &f : spec)
    for (size_t i = 0; i < f.cnt; i++) {
      const double v = flat[f.off + i];
      if (f.kind == FLOAT) {
        put_varint(out, ((uint64_t)f.number << 3) | 5);
        const float x = (float)v;
        char b[4];
        std::memcpy(b, &x, 4);  // little-endian host
        out.append(b, 4);
      } else {
        put_varint(out, ((uint64_t)f.number << 3) | 0);
        put_varint(out, (uint64_t)(v < 0 ? 0 : v + 0.5));
      }
    }
  return out;
}

// fills the fields present in the message; repeated fields must arrive with exactly their element count
inline void decode(const std::vector<field_spec> &spec, const std::string &buf, std::vector<double> &flat) {
  std::vector<size_t> seen(spec.size(), 0);
  size_t pos = 0;
  while (pos < buf.size()) {
    const uint64_t key = get_varint(buf, pos);
    const int num = (int)(key >> 3), wt = (int)(key & 7);
    std::string raw;
    uint64_t vint = 0;
    if (wt == 0) vint = get_varint(buf, pos);
    else if (wt == 5 || wt == 1) {
      const size_t n = wt == 5 ? 4 : 8;
      if (pos + n > buf.size()) throw std::runtime_error("truncated message");
      raw.assign(buf, pos, n); pos += n;
    } else if (wt == 2) {
      const uint64_t n = get_varint(buf, pos);
      if (pos + n > buf.size()) throw std::runtime_error("truncated message");
      raw.assign(buf, pos, (size_t)n); pos += (size_t)n;
    } else throw std::runtime_error("unsupported wire type");
    size_t fi = 0;
    for (; fi < spec.size(); fi++) if (spec[fi].number == num) break;
    if (fi == spec.size()) continue;  // unknown field: skipped, as protobuf does
    const field_spec &f = spec[fi];
    auto store = [&](double v) {
      if (f.repeated) {
        if (seen[fi] >= f.cnt) throw std::runtime_error("wrong dimension");
        flat[f.off + seen[fi]++] = v;
      } else { flat[f.off] = v; seen[fi] = 1; }
    };
    if (f.kind == FLOAT) {
      if (wt == 5) { float x; std::memcpy(&x, raw.data(), 4); store(x); }
      else if (wt == 2) { for (size_t p = 0; p + 4 <= raw.size(); p += 4) { float x; std::memcpy(&x, raw.data() + p, 4); store(x); } }  // packed
      else throw std::runtime_error("wire type does not match the field");
    } else {
      if (wt == 0) store((double)vint);
      else if (wt == 2) { size_t p = 0; while (p < raw.size()) store((double)get_varint(raw, p)); }  // packed
      else throw std::runtime_error("wire type does not match the field");
    }
  }
  for (size_t fi = 0; fi < spec.size(); fi++)
    if (spec[fi].repeated && seen[fi] != 0 && seen[fi] != spec[fi].cnt) throw std::runtime_error("wrong dimension");
}

}  // namespace wire
}  // namespace b200
}  // namespace models
}  // namespace microscopes
#endif
