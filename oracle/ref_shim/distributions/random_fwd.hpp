// Shim for the one header of the un-vendored `distributions` library that the
// reference's plugin boundary needs (include/microscopes/common/random_fwd.hpp:2-5).
// The reference's Cython layer states the same identity: microscopes/common/_random_fwd_h.pxd:1-8.
#pragma once
#include <random>
namespace distributions {
typedef std::default_random_engine rng_t;
}
