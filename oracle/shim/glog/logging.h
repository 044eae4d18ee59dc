// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for google-glog, which the
// reference links (voxblox/package.xml) but which is absent from this image.
// CHECK* abort with a message (glog semantics), DCHECK* compile to nothing
// (release build), LOG/VLOG swallow their stream.
#ifndef VBX_ORACLE_SHIM_GLOG_LOGGING_H_
#define VBX_ORACLE_SHIM_GLOG_LOGGING_H_

#include <cstdlib>
#include <iostream>
#include <sstream>

namespace vbx_shim_glog {
struct NullStream {
  template <typename T>
  NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream ss;
  FatalStream(const char* file, int line, const char* what) {
    ss << "[FATAL " << file << ":" << line << "] " << what << " ";
  }
  template <typename T>
  FatalStream& operator<<(const T& v) { ss << v; return *this; }
  FatalStream& operator<<(std::ostream& (*f)(std::ostream&)) { ss << f; return *this; }
  [[noreturn]] ~FatalStream() {
    std::cerr << ss.str() << std::endl;
    std::abort();
  }
};
struct Voidify {
  void operator&(const NullStream&) {}
  void operator&(const FatalStream&) {}
};
template <typename T>
T CheckNotNull(const char* file, int line, const char* names, T&& t) {
  if (t == nullptr) { FatalStream(file, line, names) << "must be non NULL"; }
  return std::forward<T>(t);
}
}  // namespace vbx_shim_glog

#define VBX_SHIM_NULL ::vbx_shim_glog::NullStream()
#define VBX_SHIM_FATAL(what) ::vbx_shim_glog::FatalStream(__FILE__, __LINE__, what)

#define CHECK(cond) \
  (cond) ? (void)0 : ::vbx_shim_glog::Voidify() & VBX_SHIM_FATAL("Check failed: " #cond)
#define VBX_SHIM_CHECK_OP(a, b, op) \
  ((a)op(b)) ? (void)0 : ::vbx_shim_glog::Voidify() & VBX_SHIM_FATAL("Check failed: " #a " " #op " " #b)
#define CHECK_EQ(a, b) VBX_SHIM_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) VBX_SHIM_CHECK_OP(a, b, !=)
#define CHECK_LT(a, b) VBX_SHIM_CHECK_OP(a, b, <)
#define CHECK_LE(a, b) VBX_SHIM_CHECK_OP(a, b, <=)
#define CHECK_GT(a, b) VBX_SHIM_CHECK_OP(a, b, >)
#define CHECK_GE(a, b) VBX_SHIM_CHECK_OP(a, b, >=)
#define CHECK_NEAR(a, b, tol)                                     \
  (((a) - (b) <= (tol)) && ((b) - (a) <= (tol)))                  \
      ? (void)0                                                   \
      : ::vbx_shim_glog::Voidify() & VBX_SHIM_FATAL("Check failed: " #a " near " #b)
#define CHECK_NOTNULL(p) \
  ::vbx_shim_glog::CheckNotNull(__FILE__, __LINE__, "'" #p "'", (p))

#define VBX_SHIM_DNULL(cond) \
  true ? (void)0 : ::vbx_shim_glog::Voidify() & VBX_SHIM_NULL
#define DCHECK(cond) VBX_SHIM_DNULL(cond)
#define DCHECK_EQ(a, b) VBX_SHIM_DNULL(0)
#define DCHECK_NE(a, b) VBX_SHIM_DNULL(0)
#define DCHECK_LT(a, b) VBX_SHIM_DNULL(0)
#define DCHECK_LE(a, b) VBX_SHIM_DNULL(0)
#define DCHECK_GT(a, b) VBX_SHIM_DNULL(0)
#define DCHECK_GE(a, b) VBX_SHIM_DNULL(0)
#define DCHECK_NOTNULL(p) (p)

#define VBX_SHIM_LOG_INFO VBX_SHIM_NULL
#define VBX_SHIM_LOG_WARNING VBX_SHIM_NULL
#define VBX_SHIM_LOG_ERROR VBX_SHIM_NULL
#define VBX_SHIM_LOG_FATAL VBX_SHIM_FATAL("LOG(FATAL)")
#define LOG(sev) VBX_SHIM_LOG_##sev
#define LOG_IF(sev, cond) (!(cond)) ? (void)0 : ::vbx_shim_glog::Voidify() & LOG(sev)
#define LOG_FIRST_N(sev, n) LOG(sev)
#define LOG_EVERY_N(sev, n) LOG(sev)
#define VLOG(n) VBX_SHIM_NULL
#define VLOG_IS_ON(n) false

#endif  // VBX_ORACLE_SHIM_GLOG_LOGGING_H_
