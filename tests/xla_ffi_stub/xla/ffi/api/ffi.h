// TEST STUB - not XLA, never shipped, never linked.  jaxlib is not installable in this image, so the XLA-FFI adapter
// (gvi_gaussian_process_b200/csrc/ffi/gvigp_xla_ffi.cc) cannot be compiled against the real header.  This file models
// only the part of the public xla/ffi/api/ffi.h surface the adapter uses (Buffer / ResultBuffer accessors, Error, the
// Ffi::Bind() builder and XLA_FFI_DEFINE_HANDLER_SYMBOL), written from the published API, with ONE purpose:
// tests/test_abi.py compiles the adapter against it with `g++ -fsyntax-only`, which type-checks
//   (1) every forwarded call against the C ABI of include/gvigp.h (argument count, order and types), and
//   (2) every handler's C++ signature against its binding (context, arguments, attributes and results in chain order).
// It says nothing about XLA's runtime behaviour.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla {
namespace ffi {

enum class DataType : uint8_t { U8, S64, F64 };
inline constexpr DataType U8 = DataType::U8;
inline constexpr DataType S64 = DataType::S64;
inline constexpr DataType F64 = DataType::F64;

template <DataType T> struct NativeType;
template <> struct NativeType<DataType::U8> { using type = uint8_t; };
template <> struct NativeType<DataType::S64> { using type = int64_t; };
template <> struct NativeType<DataType::F64> { using type = double; };

template <typename T> struct Span {
  const T* ptr = nullptr;
  size_t len = 0;
  size_t size() const { return len; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

template <DataType T> struct Buffer {
  using Native = typename NativeType<T>::type;
  Native* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  Span<int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
};

template <typename T> struct Result {
  T value;
  T* operator->() { return &value; }
  T& operator*() { return value; }
};
template <DataType T> using ResultBuffer = Result<Buffer<T>>;

struct Error {
  static Error Success() { return {}; }
  static Error Internal(std::string) { return {}; }
  static Error InvalidArgument(std::string) { return {}; }
};

template <typename T> struct PlatformStream {};

namespace stub_detail {
template <typename T> struct CtxType;
template <typename S> struct CtxType<PlatformStream<S>> { using type = S; };
}  // namespace stub_detail

// Accumulates the handler's parameter list in chain order, as the real binding does.
template <typename... Ts> struct Binding {
  template <typename T> constexpr auto Ctx() const { return Binding<Ts..., typename stub_detail::CtxType<T>::type>{}; }
  template <typename T> constexpr auto Arg() const { return Binding<Ts..., T>{}; }
  template <typename T> constexpr auto Ret() const { return Binding<Ts..., Result<T>>{}; }
  template <typename T> constexpr auto Attr(const char*) const { return Binding<Ts..., T>{}; }
  template <typename F> static constexpr bool matches = std::is_same_v<F, Error (*)(Ts...)>;
};

struct Ffi {
  static constexpr Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                              \
  static_assert(decltype(binding)::template matches<decltype(&impl)>,                                  \
                #sym ": the handler's signature does not match its binding (ctx, args, attrs, rets)"); \
  extern "C" void* sym() { return nullptr; }
