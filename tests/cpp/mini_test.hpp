// Tiny test harness for the C++ shim tests (Catch2 is not available in this image).
#pragma once
#include <cmath>
#include <cstdio>
#include <exception>
#include <functional>
#include <string>
#include <vector>

namespace mini {
struct Case { std::string name; std::function<void()> fn; };
inline std::vector<Case> &cases() { static std::vector<Case> c; return c; }
inline int &failures() { static int f = 0; return f; }
inline int &checks() { static int c = 0; return c; }
struct Reg { Reg(const char *n, std::function<void()> f) { cases().push_back({n, std::move(f)}); } };
inline int run_all() {
    int failed_cases = 0;
    for (auto &c : cases()) {
        int before = failures();
        try { c.fn(); } catch (std::exception const &e) {
            std::printf("  EXCEPTION in %s: %s\n", c.name.c_str(), e.what());
            failures()++;
        }
        bool ok = failures() == before;
        std::printf("[%s] %s\n", ok ? " OK " : "FAIL", c.name.c_str());
        failed_cases += !ok;
    }
    std::printf("%d cases, %d checks, %d failed checks\n", (int)cases().size(), checks(), failures());
    return failed_cases ? 1 : 0;
}
} // namespace mini

#define TEST_CASE(name) static void name(); static mini::Reg reg_##name(#name, name); static void name()
#define CHECK(cond) do { mini::checks()++; if (!(cond)) { mini::failures()++; std::printf("  CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); } } while (0)
#define CHECK_CLOSE(a, b, tol) do { mini::checks()++; double a_ = std::abs((a) - (b)); if (!(a_ <= (tol))) { mini::failures()++; \
    std::printf("  CHECK_CLOSE failed %s:%d: |%s - %s| = %g > %g\n", __FILE__, __LINE__, #a, #b, a_, (double)(tol)); } } while (0)
#define CHECK_THROWS_AS(expr, Ex) do { mini::checks()++; bool ok_ = false; try { expr; } catch (Ex const &) { ok_ = true; } catch (...) {} \
    if (!ok_) { mini::failures()++; std::printf("  CHECK_THROWS_AS failed %s:%d: %s\n", __FILE__, __LINE__, #expr); } } while (0)
