// -*- mode: c++ -*-
// catch2/catch.hpp -- a minimal stand-in for the part of Catch2 v2 (2.9.2, /root/reference/conanfile.txt:2)
// that the reference's own test sources use, so that /root/reference/tests/*.cpp compile UNCHANGED and run
// against liblambrex.so (Catch2 is not installed in this image and there is no network).  TEST
// INFRASTRUCTURE, written from scratch; not a copy of Catch2.
//
// Supported: TEST_CASE(name, tags), SECTION (nested; a test case is re-run once per leaf section, like Catch2),
// REQUIRE / REQUIRE_FALSE / CHECK / CHECK_FALSE, INFO, Approx (epsilon = 100 * FLT_EPSILON, margin 0, scale 0 --
// Catch2 v2's defaults -- with .epsilon()/.margin()/.scale()), CATCH_CONFIG_MAIN, Catch::TestEventListenerBase
// with testRunStarting / testRunEnded, CATCH_REGISTER_LISTENER.
//
// Command line of a test binary:  [name-substring] [--list-tests]
// Environment: CATCH_SHIM_CONTINUE=1 turns REQUIRE into CHECK (count every failing assertion instead of aborting
// the test case at the first one) -- used to count how many points of `ml_pulse Regression` agree;
// CATCH_SHIM_MAX_PRINT=n reports the first n failing assertions in full (default 20).
// Last line of output:  "catch-shim: test cases: N | passed P | failed F | assertions: A | failed assertions: X".
#ifndef LBX_CATCH_SHIM_HPP
#define LBX_CATCH_SHIM_HPP

#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace Catch {

// ---------------------------------------------------------------------------------------- Approx
class Approx {
 public:
  explicit Approx(double v) : v_(v) {}
  Approx& epsilon(double e) { eps_ = e; return *this; }
  Approx& margin(double m) { margin_ = m; return *this; }
  Approx& scale(double s) { scale_ = s; return *this; }
  Approx operator-() const { Approx a(-v_); a.eps_ = eps_; a.margin_ = margin_; a.scale_ = scale_; return a; }
  bool equals(double x) const {
    auto within = [](double a, double b, double m) { return (a + m >= b) && (b + m >= a); };
    return within(x, v_, margin_) || within(x, v_, eps_ * (scale_ + std::fabs(std::isinf(v_) ? 0.0 : v_)));
  }
  double value() const { return v_; }

 private:
  double v_, eps_ = static_cast<double>(FLT_EPSILON) * 100.0, margin_ = 0.0, scale_ = 0.0;
};
template <class T> bool operator==(const T& x, const Approx& a) { return a.equals(static_cast<double>(x)); }
template <class T> bool operator==(const Approx& a, const T& x) { return a.equals(static_cast<double>(x)); }
template <class T> bool operator!=(const T& x, const Approx& a) { return !a.equals(static_cast<double>(x)); }
template <class T> bool operator!=(const Approx& a, const T& x) { return !a.equals(static_cast<double>(x)); }
inline std::ostream& operator<<(std::ostream& os, const Approx& a) { return os << "Approx(" << a.value() << ")"; }
namespace Detail { using Catch::Approx; }

// ---------------------------------------------------------------------------------------- listener
struct TestRunInfo { std::string name; };
struct Totals { int testCases = 0, failedCases = 0; long assertions = 0, failedAssertions = 0; };
struct TestRunStats { TestRunInfo runInfo; Totals totals; bool aborting = false; };
struct ReporterConfig {};
struct TestEventListenerBase {
  explicit TestEventListenerBase(ReporterConfig const& = ReporterConfig()) {}
  virtual ~TestEventListenerBase() = default;
  virtual void testRunStarting(TestRunInfo const&) {}
  virtual void testRunEnded(TestRunStats const&) {}
};

// ---------------------------------------------------------------------------------------- registry + sections
struct TestFailure : std::exception {
  const char* what() const noexcept override { return "REQUIRE failed"; }
};

struct SectionNode {
  std::string name;
  SectionNode* parent = nullptr;
  std::vector<std::unique_ptr<SectionNode>> children;
  bool done = false;
  bool child_entered_this_run = false;
  SectionNode* child(const std::string& n) {
    for (auto& c : children)
      if (c->name == n) return c.get();
    children.emplace_back(new SectionNode);
    children.back()->name = n;
    children.back()->parent = this;
    return children.back().get();
  }
  bool all_children_done() const {
    for (auto& c : children)
      if (!c->done) return false;
    return true;
  }
};

struct TestCase { std::string name, tags; void (*fn)(); };

struct Registry {
  std::vector<TestCase> cases;
  std::vector<std::unique_ptr<TestEventListenerBase>> listeners;
  SectionNode* cur = nullptr;              // innermost open section of the running test case
  std::vector<std::string> infos;          // active INFO messages
  Totals totals;
  bool soft_require = false;
  bool case_failed = false;
  long max_print = 20;                     // failing assertions reported in full (CATCH_SHIM_MAX_PRINT)
  static Registry& get() { static Registry r; return r; }
};

struct AutoReg {
  AutoReg(const char* name, const char* tags, void (*fn)()) { Registry::get().cases.push_back({name, tags, fn}); }
};
template <class L>
struct ListenerReg {
  ListenerReg() { Registry::get().listeners.emplace_back(new L(ReporterConfig())); }
};

class Section {
 public:
  explicit Section(const std::string& name) {
    Registry& r = Registry::get();
    SectionNode* parent = r.cur;
    node_ = parent->child(name);
    if (!node_->done && !parent->child_entered_this_run) {
      parent->child_entered_this_run = true;
      node_->child_entered_this_run = false;
      r.cur = node_;
      entered_ = true;
    }
  }
  ~Section() {
    if (!entered_) return;
    node_->done = node_->all_children_done();
    Registry::get().cur = node_->parent;
  }
  explicit operator bool() const { return entered_; }

 private:
  SectionNode* node_ = nullptr;
  bool entered_ = false;
};

struct ScopedInfo {
  explicit ScopedInfo(const std::string& s) { Registry::get().infos.push_back(s); }
  ~ScopedInfo() { Registry::get().infos.pop_back(); }
};

inline void assertion(bool ok, bool fatal, const char* macro, const char* expr, const char* file, int line) {
  Registry& r = Registry::get();
  ++r.totals.assertions;
  if (ok) return;
  ++r.totals.failedAssertions;
  r.case_failed = true;
  if (r.totals.failedAssertions <= r.max_print) {
    std::cout << file << ":" << line << ": FAILED: " << macro << "( " << expr << " )\n";
    for (const auto& s : r.infos) std::cout << "  with message: " << s << "\n";
  }
  if (fatal && !r.soft_require) throw TestFailure();
}

inline int run(int argc, char** argv) {
  Registry& r = Registry::get();
  std::string filter;
  bool list = false;
  for (int a = 1; a < argc; ++a) {
    if (!std::strcmp(argv[a], "--list-tests")) list = true;
    else if (argv[a][0] != '-') filter = argv[a];
  }
  const char* soft = std::getenv("CATCH_SHIM_CONTINUE");
  r.soft_require = soft && soft[0] && soft[0] != '0';
  if (const char* mp = std::getenv("CATCH_SHIM_MAX_PRINT")) r.max_print = std::atol(mp);
  if (list) {
    for (const auto& c : r.cases) std::cout << c.name << "  " << c.tags << "\n";
    return 0;
  }
  TestRunStats stats;
  stats.runInfo.name = argc > 0 ? argv[0] : "tests";
  for (auto& l : r.listeners) l->testRunStarting(stats.runInfo);
  for (const auto& c : r.cases) {
    if (!filter.empty() && c.name.find(filter) == std::string::npos) continue;
    ++r.totals.testCases;
    r.case_failed = false;
    const long a0 = r.totals.assertions, f0 = r.totals.failedAssertions;
    SectionNode root;
    root.name = c.name;
    int runs = 0;
    do {                                        // once per leaf section
      root.child_entered_this_run = false;
      r.cur = &root;
      r.infos.clear();
      try {
        c.fn();
      } catch (const TestFailure&) {
      } catch (const std::exception& e) {
        r.case_failed = true;
        std::cout << c.name << ": unexpected exception: " << e.what() << "\n";
      } catch (...) {
        r.case_failed = true;
        std::cout << c.name << ": unexpected exception\n";
      }
      ++runs;
    } while (!root.all_children_done() && runs < 10000);
    if (r.case_failed) ++r.totals.failedCases;
    std::cout << (r.case_failed ? "[FAILED] " : "[  OK  ] ") << c.name << "  (" << runs << " run" << (runs == 1 ? "" : "s") << ", "
              << (r.totals.assertions - a0) << " assertions, " << (r.totals.failedAssertions - f0) << " failed)" << std::endl;
  }
  stats.totals = r.totals;
  for (auto& l : r.listeners) l->testRunEnded(stats);
  std::cout << "catch-shim: test cases: " << r.totals.testCases << " | passed " << (r.totals.testCases - r.totals.failedCases)
            << " | failed " << r.totals.failedCases << " | assertions: " << r.totals.assertions
            << " | failed assertions: " << r.totals.failedAssertions << std::endl;
  return r.totals.failedCases > 255 ? 255 : r.totals.failedCases;
}

}  // namespace Catch

using Catch::Approx;

#define LBX_CATCH_CAT2(a, b) a##b
#define LBX_CATCH_CAT(a, b) LBX_CATCH_CAT2(a, b)
#define LBX_CATCH_UNIQUE(p) LBX_CATCH_CAT(p, __LINE__)

#define TEST_CASE(...)                                                                               \
  static void LBX_CATCH_UNIQUE(lbx_catch_test_)();                                                    \
  namespace { Catch::AutoReg LBX_CATCH_UNIQUE(lbx_catch_reg_)(LBX_CATCH_NAME(__VA_ARGS__, ""), LBX_CATCH_TAGS(__VA_ARGS__, "", ""), \
                                                              &LBX_CATCH_UNIQUE(lbx_catch_test_)); }  \
  static void LBX_CATCH_UNIQUE(lbx_catch_test_)()
#define LBX_CATCH_NAME(n, ...) n
#define LBX_CATCH_TAGS(n, t, ...) t

#define SECTION(...) if (Catch::Section LBX_CATCH_UNIQUE(lbx_catch_section_){LBX_CATCH_NAME(__VA_ARGS__, "")})
#define INFO(msg)                                         \
  std::ostringstream LBX_CATCH_UNIQUE(lbx_catch_os_);     \
  LBX_CATCH_UNIQUE(lbx_catch_os_) << msg;                 \
  Catch::ScopedInfo LBX_CATCH_UNIQUE(lbx_catch_info_)(LBX_CATCH_UNIQUE(lbx_catch_os_).str());
#define CAPTURE(x) INFO(#x " := " << (x))
#define REQUIRE(...) Catch::assertion(static_cast<bool>(__VA_ARGS__), true, "REQUIRE", #__VA_ARGS__, __FILE__, __LINE__)
#define REQUIRE_FALSE(...) Catch::assertion(!static_cast<bool>(__VA_ARGS__), true, "REQUIRE_FALSE", #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK(...) Catch::assertion(static_cast<bool>(__VA_ARGS__), false, "CHECK", #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK_FALSE(...) Catch::assertion(!static_cast<bool>(__VA_ARGS__), false, "CHECK_FALSE", #__VA_ARGS__, __FILE__, __LINE__)
#define CATCH_REGISTER_LISTENER(L) namespace { Catch::ListenerReg<L> LBX_CATCH_UNIQUE(lbx_catch_listener_); }

#ifdef CATCH_CONFIG_MAIN
int main(int argc, char** argv) { return Catch::run(argc, argv); }
#endif

#endif
