// STAND-IN for muu/preprocessor.h (muu is not vendored with marzer/rt; see oracle/ref_shim/README.md).
// Only what the reference's renderer translation units use.  TEST INFRASTRUCTURE.
#pragma once
#define MUU_GCC 1
#define MUU_CLANG 0
#define MUU_DISABLE_WARNINGS static_assert(true)
#define MUU_ENABLE_WARNINGS static_assert(true)
#define MUU_DISABLE_SPAM_WARNINGS static_assert(true)
#define MUU_PUSH_WARNINGS static_assert(true)
#define MUU_POP_WARNINGS static_assert(true)
#define MUU_FORCE_NDEBUG_OPTIMIZATIONS static_assert(true)
#define MUU_PURE_INLINE_GETTER [[nodiscard]] inline
#define MUU_PURE_GETTER [[nodiscard]]
#define MUU_PURE
#define MUU_NODISCARD [[nodiscard]]
#define MUU_NODISCARD_CTOR
#define MUU_ALWAYS_INLINE inline
#define MUU_VECTORCALL
#define MUU_TRIVIAL_ABI
#define MUU_ABSTRACT_INTERFACE
#define MUU_CONSTEVAL consteval
#define MUU_ATTR(...) __attribute__((__VA_ARGS__))
#define MUU_FMA_BLOCK static_assert(true)
#define MUU_UNLIKELY(...) (__builtin_expect(!!(__VA_ARGS__), 0))
#define MUU_ASSUME(...) static_cast<void>(0)
#define MUU_CONSTEXPR_SAFE_ASSERT(...) static_cast<void>(0)
#define MUU_CONCAT_2(a, b) a##b
#define MUU_CONCAT(a, b) MUU_CONCAT_2(a, b)
#define MUU_MAKE_STRING_2(s) #s
#define MUU_MAKE_STRING(s) MUU_MAKE_STRING_2(s)
#define MUU_CONSTRAINED_TEMPLATE(cond, ...) template <__VA_ARGS__> requires(cond)
