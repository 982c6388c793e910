// plugin_api.hpp -- the plugin boundary the adapters in gpu_models.hpp implement.
//
// When the reference tree is on the include path (-DMSB_USE_REFERENCE_HEADERS
// -I<reference>/include) this header simply includes the reference's own
// <microscopes/models/base.hpp>, and the adapters are a drop-in for
// distributions_model<T> (include/microscopes/models/distributions.hpp:395-509).
//
// Without it (the GPU box has no reference tree) the block below declares the
// same interface -- same namespaces, class names, method names, argument order
// and meaning as include/microscopes/models/base.hpp:21-62,
// common/runtime_type.hpp:65-141 and common/runtime_value.hpp:9-98 -- so that
// the same adapter and test sources compile unchanged.
#pragma once

#ifdef MSB_USE_REFERENCE_HEADERS
#include <microscopes/models/base.hpp>
#else

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

// primitive type ids: include/microscopes/common/type_info.h:10-34 (same order, same values)
enum primitive_type {
  TYPE_B, TYPE_I8, TYPE_U8, TYPE_I16, TYPE_U16, TYPE_I32, TYPE_U32, TYPE_I64, TYPE_U64, TYPE_F32, TYPE_F64,
  TYPE_NELEMS
};

namespace microscopes {
namespace common {

typedef std::default_random_engine rng_t;  // random_fwd.hpp:5 / _random_fwd_h.pxd:1-8
typedef std::string hyperparam_bag_t;      // typedefs.hpp:13-15
typedef std::string suffstats_bag_t;

namespace detail {
inline unsigned prim_size(primitive_type t) {
  static const unsigned sizes[TYPE_NELEMS] = {1, 1, 1, 2, 2, 4, 4, 8, 8, 4, 8};
  return sizes[t];
}
template <typename T> struct prim_of;
template <> struct prim_of<bool> { static const primitive_type value = TYPE_B; };
template <> struct prim_of<int8_t> { static const primitive_type value = TYPE_I8; };
template <> struct prim_of<uint8_t> { static const primitive_type value = TYPE_U8; };
template <> struct prim_of<int16_t> { static const primitive_type value = TYPE_I16; };
template <> struct prim_of<uint16_t> { static const primitive_type value = TYPE_U16; };
template <> struct prim_of<int32_t> { static const primitive_type value = TYPE_I32; };
template <> struct prim_of<uint32_t> { static const primitive_type value = TYPE_U32; };
template <> struct prim_of<int64_t> { static const primitive_type value = TYPE_I64; };
template <> struct prim_of<uint64_t> { static const primitive_type value = TYPE_U64; };
template <> struct prim_of<float> { static const primitive_type value = TYPE_F32; };
template <> struct prim_of<double> { static const primitive_type value = TYPE_F64; };

// what runtime_cast::cast / uncast do (runtime_type.hpp:145-196), through a double
inline double load_as_double(const uint8_t *p, primitive_type t) {
  switch (t) {
    case TYPE_B: return *reinterpret_cast<const bool *>(p);
    case TYPE_I8: return *reinterpret_cast<const int8_t *>(p);
    case TYPE_U8: return *p;
    case TYPE_I16: { int16_t v; std::memcpy(&v, p, 2); return v; }
    case TYPE_U16: { uint16_t v; std::memcpy(&v, p, 2); return v; }
    case TYPE_I32: { int32_t v; std::memcpy(&v, p, 4); return v; }
    case TYPE_U32: { uint32_t v; std::memcpy(&v, p, 4); return v; }
    case TYPE_I64: { int64_t v; std::memcpy(&v, p, 8); return (double)v; }
    case TYPE_U64: { uint64_t v; std::memcpy(&v, p, 8); return (double)v; }
    case TYPE_F32: { float v; std::memcpy(&v, p, 4); return v; }
    default: { double v; std::memcpy(&v, p, 8); return v; }
  }
}
inline void store_from_double(uint8_t *p, primitive_type t, double x) {
  switch (t) {
    case TYPE_B: *reinterpret_cast<bool *>(p) = x != 0; break;
    case TYPE_I8: *reinterpret_cast<int8_t *>(p) = (int8_t)x; break;
    case TYPE_U8: *p = (uint8_t)x; break;
    case TYPE_I16: { int16_t v = (int16_t)x; std::memcpy(p, &v, 2); } break;
    case TYPE_U16: { uint16_t v = (uint16_t)x; std::memcpy(p, &v, 2); } break;
    case TYPE_I32: { int32_t v = (int32_t)x; std::memcpy(p, &v, 4); } break;
    case TYPE_U32: { uint32_t v = (uint32_t)x; std::memcpy(p, &v, 4); } break;
    case TYPE_I64: { int64_t v = (int64_t)x; std::memcpy(p, &v, 8); } break;
    case TYPE_U64: { uint64_t v = (uint64_t)x; std::memcpy(p, &v, 8); } break;
    case TYPE_F32: { float v = (float)x; std::memcpy(p, &v, 4); } break;
    default: std::memcpy(p, &x, 8); break;
  }
}
}  // namespace detail

class runtime_type {
public:
  runtime_type() : t_(), n_(), vec_() {}
  runtime_type(primitive_type t) : t_(t), n_(1), vec_(false) {}
  runtime_type(primitive_type t, unsigned n) : t_(t), n_(n), vec_(true) {}
  primitive_type t() const { return t_; }
  unsigned psize() const { return detail::prim_size(t_); }
  unsigned size() const { return n_ * psize(); }
  unsigned n() const { return n_; }
  bool vec() const { return vec_; }
  bool operator==(const runtime_type &o) const { return t_ == o.t_ && n_ == o.n_ && vec_ == o.vec_; }
  bool operator!=(const runtime_type &o) const { return !(*this == o); }

private:
  primitive_type t_;
  unsigned n_;
  bool vec_;
};

class value_accessor {
public:
  value_accessor() : data_(), mask_(), type_() {}
  template <typename T>
  value_accessor(const T *data)
      : data_(reinterpret_cast<const uint8_t *>(data)), mask_(nullptr), type_(runtime_type(detail::prim_of<T>::value)) {}
  value_accessor(const uint8_t *data, const bool *mask, const runtime_type &type) : data_(data), mask_(mask), type_(type) {}
  const runtime_type &type() const { return type_; }
  unsigned shape() const { return type_.n(); }
  bool ismasked(size_t idx) const { return mask_ ? mask_[idx] : false; }
  bool anymasked() const {
    if (!mask_) return false;
    for (size_t i = 0; i < shape(); i++) if (mask_[i]) return true;
    return false;
  }
  template <typename T> T get(size_t idx = 0) const {
    return (T)detail::load_as_double(data_ + idx * type_.psize(), type_.t());
  }

private:
  const uint8_t *data_;
  const bool *mask_;
  runtime_type type_;
};

class value_mutator {
public:
  value_mutator() : data_(), type_() {}
  template <typename T>
  value_mutator(T *data) : data_(reinterpret_cast<uint8_t *>(data)), type_(runtime_type(detail::prim_of<T>::value)) {}
  value_mutator(uint8_t *data, const runtime_type &type) : data_(data), type_(type) {}
  const runtime_type &type() const { return type_; }
  unsigned shape() const { return type_.n(); }
  template <typename T> void set(T t, size_t idx = 0) {
    detail::store_from_double(data_ + idx * type_.psize(), type_.t(), (double)t);
  }
  value_accessor accessor() const { return value_accessor(data_, nullptr, type_); }

private:
  uint8_t *data_;
  runtime_type type_;
};

}  // namespace common

namespace models {

class hypers;

class group {
public:
  virtual ~group() {}
  virtual void add_value(const hypers &m, const common::value_accessor &value, common::rng_t &rng) = 0;
  virtual void remove_value(const hypers &m, const common::value_accessor &value, common::rng_t &rng) = 0;
  virtual float score_value(const hypers &m, const common::value_accessor &value, common::rng_t &rng) const = 0;
  virtual float score_data(const hypers &m, common::rng_t &rng) const = 0;
  virtual void sample_value(const hypers &m, common::value_mutator &value, common::rng_t &rng) const = 0;
  virtual common::suffstats_bag_t get_ss() const = 0;
  virtual void set_ss(const common::suffstats_bag_t &ss) = 0;
  virtual void set_ss(const group &g) = 0;
  virtual common::value_mutator get_ss_mutator(const std::string &key) = 0;
  virtual std::string debug_str() const = 0;
};

class hypers {
public:
  virtual ~hypers() {}
  virtual common::hyperparam_bag_t get_hp() const = 0;
  virtual void set_hp(const common::hyperparam_bag_t &hp) = 0;
  virtual void set_hp(const hypers &s) = 0;
  virtual common::value_mutator get_hp_mutator(const std::string &key) = 0;
  virtual std::shared_ptr<group> create_group(common::rng_t &rng) const = 0;
  virtual std::string debug_str() const = 0;
};

class model {
public:
  virtual ~model() {}
  virtual std::shared_ptr<hypers> create_hypers() const = 0;
  virtual common::runtime_type get_runtime_type() const = 0;
};

}  // namespace models
}  // namespace microscopes

#endif  // MSB_USE_REFERENCE_HEADERS
