// Pre-include used ONLY to compile the unmodified reference kNN extension
// (/root/reference/torch_knnquery/src/knnquery.cu) against torch 2.11 headers.
// The reference calls AT_DISPATCH_FLOATING_TYPES(x.type(), ...) (knnquery.cu:338,
// 420, 461, 494, 544); torch >= 2.x dropped the DeprecatedTypeProperties overload
// of ::detail::scalar_type that this relied on.  We add it back here so the
// reference source compiles in place, byte-for-byte unmodified.
// TEST INFRASTRUCTURE ONLY - never linked into the product library.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) {
  return t.scalarType();
}
}  // namespace detail
