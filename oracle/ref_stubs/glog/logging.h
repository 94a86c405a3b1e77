// Stand-in for <glog/logging.h> (TEST INFRASTRUCTURE, see oracle/ref_stubs/README.md): CHECK* abort with a
// message, LOG(ERROR/WARNING/INFO) print to stderr, LOG(FATAL) aborts, DCHECK* compile but never evaluate
// (the reference is built in Release, catkin default in its README, where glog's DCHECKs vanish).
#pragma once
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>

namespace stub_glog {
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };
class Message {
 public:
  Message(const char* file, int line, int severity) : severity_(severity) { s_ << file << ":" << line << "] "; }
  ~Message() noexcept(false) {
    std::cerr << "IWEF"[severity_] << " " << s_.str() << std::endl;
    if (severity_ == FATAL) std::abort();
  }
  std::ostream& stream() { return s_; }

 private:
  int severity_;
  std::ostringstream s_;
};
struct Voidify {
  void operator&(std::ostream&) {}
};
template <class T>
T&& CheckNotNull(const char* file, int line, const char* what, T&& t) {
  if (t == nullptr) Message(file, line, FATAL).stream() << what;
  return std::forward<T>(t);
}
}  // namespace stub_glog

#define LOG(sev) stub_glog::Message(__FILE__, __LINE__, stub_glog::sev).stream()
#define VLOG(n) (true) ? (void)0 : stub_glog::Voidify() & LOG(INFO)
#define LOG_IF(sev, cond) !(cond) ? (void)0 : stub_glog::Voidify() & LOG(sev)
#define CHECK(cond) (cond) ? (void)0 : stub_glog::Voidify() & LOG(FATAL) << "Check failed: " #cond " "
#define STUB_CHECK_OP(a, op, b) CHECK((a)op(b))
#define CHECK_EQ(a, b) STUB_CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) STUB_CHECK_OP(a, !=, b)
#define CHECK_LT(a, b) STUB_CHECK_OP(a, <, b)
#define CHECK_LE(a, b) STUB_CHECK_OP(a, <=, b)
#define CHECK_GT(a, b) STUB_CHECK_OP(a, >, b)
#define CHECK_GE(a, b) STUB_CHECK_OP(a, >=, b)
#define CHECK_NEAR(a, b, tol) CHECK(std::abs((a) - (b)) <= (tol))
#define CHECK_NOTNULL(p) stub_glog::CheckNotNull(__FILE__, __LINE__, "'" #p "' Must be non NULL", (p))
#define DCHECK(cond) while (false) CHECK(cond)
#define DCHECK_EQ(a, b) while (false) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) while (false) CHECK_NE(a, b)
#define DCHECK_LT(a, b) while (false) CHECK_LT(a, b)
#define DCHECK_LE(a, b) while (false) CHECK_LE(a, b)
#define DCHECK_GT(a, b) while (false) CHECK_GT(a, b)
#define DCHECK_GE(a, b) while (false) CHECK_GE(a, b)
#define DCHECK_NOTNULL(p) (p)
