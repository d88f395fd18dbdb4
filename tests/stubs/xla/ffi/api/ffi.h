// MINIMAL STAND-IN for jaxlib's `xla/ffi/api/ffi.h` -- TEST INFRASTRUCTURE ONLY.
//
// JAX / jaxlib are not installed in this image, so `csrc/ffi_shim.cc` cannot be compiled against the real
// header here.  This file declares just the part of the XLA FFI C++ API the shim uses, with the same names and
// call shapes, so that `g++ -fsyntax-only` (tests/test_abi_cpu.py) type-checks every handler against its binding:
// `Binding::To(fn)` static_asserts that `fn` is invocable with exactly the argument list the
// `.Ctx/.Arg/.Attr/.Ret/.RemainingArgs/.RemainingRets` chain describes, in that order, and returns `Error`.
// Nothing here executes; a build against the real jaxlib header is what ships (see the shim's own comment).
#ifndef BL_TEST_STUB_XLA_FFI_API_FFI_H_
#define BL_TEST_STUB_XLA_FFI_API_FFI_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;

namespace xla {
namespace ffi {

enum class DataType { INVALID, PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
enum class ErrorCode { kOk, kCancelled, kUnknown, kInvalidArgument, kInternal, kUnimplemented };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string message) { return Error(ErrorCode::kInvalidArgument, std::move(message)); }
  static Error Internal(std::string message) { return Error(ErrorCode::kInternal, std::move(message)); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* data, size_t size) : data_(data), size_(size) {}
  size_t size() const { return size_; }
  const T& operator[](size_t i) const { return data_[i]; }
  const T* begin() const { return data_; }
  const T* end() const { return data_ + size_; }
  const T& front() const { return data_[0]; }
  const T& back() const { return data_[size_ - 1]; }

 private:
  const T* data_ = nullptr;
  size_t size_ = 0;
};

class AnyBuffer {
 public:
  using Dimensions = Span<int64_t>;
  DataType element_type() const { return type_; }
  Dimensions dimensions() const { return Dimensions(dims_, rank_); }
  void* untyped_data() const { return data_; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }

 private:
  DataType type_ = DataType::INVALID;
  void* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0;
};

template <typename T>
class Result {
 public:
  T& operator*() { return value_; }
  T* operator->() { return &value_; }

 private:
  T value_;
};

template <typename T>
class ErrorOr {
 public:
  bool has_value() const { return true; }
  T& value() { return value_; }
  T& operator*() { return value_; }
  T* operator->() { return &value_; }

 private:
  T value_;
};

class RemainingArgs {
 public:
  size_t size() const { return 0; }
  bool empty() const { return true; }
  template <typename T>
  ErrorOr<T> get(size_t) const { return ErrorOr<T>(); }
};

class RemainingRets {
 public:
  size_t size() const { return 0; }
  bool empty() const { return true; }
  template <typename T>
  ErrorOr<Result<T>> get(size_t) const { return ErrorOr<Result<T>>(); }
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename... Ts>
struct TypeList {};
template <typename F, typename List>
struct Invocable;
template <typename F, typename... Ts>
struct Invocable<F, TypeList<Ts...>> : std::is_invocable_r<Error, F, Ts...> {};
struct Handler {};
}  // namespace internal

template <typename... Ts>
class Binding {
 public:
  template <typename T>
  auto Ctx() const { return CtxHelper<T>::apply(*this); }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingArgs> RemainingArgs() const { return {}; }
  Binding<Ts..., ::xla::ffi::RemainingRets> RemainingRets() const { return {}; }
  template <typename F>
  internal::Handler To(F&&) const {
    static_assert(internal::Invocable<F, internal::TypeList<Ts...>>::value,
                  "handler signature does not match its XLA FFI binding (order: Ctx, Arg, Attr, Ret as bound)");
    return {};
  }

 private:
  template <typename T>
  struct CtxHelper;
  template <typename S>
  struct CtxHelper<PlatformStream<S>> {
    static Binding<Ts..., S> apply(const Binding&) { return {}; }
  };
};

class Ffi {
 public:
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, fn, binding)                              \
  static const ::xla::ffi::internal::Handler symbol##_stub_handler = (binding).To(fn); \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*) { return nullptr; }

#endif  // BL_TEST_STUB_XLA_FFI_API_FFI_H_
