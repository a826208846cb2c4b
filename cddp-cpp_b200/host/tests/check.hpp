// tiny assertion harness (googletest is a network FetchContent dependency of the reference and unavailable here)
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                                  \
  do {                                                                               \
    ++g_checks;                                                                      \
    if (!(cond)) {                                                                   \
      ++g_failed;                                                                    \
      std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);           \
    }                                                                                \
  } while (0)
#define CHECK_NEAR(a, b, tol) CHECK(std::fabs((a) - (b)) <= (tol))
#define RUN(test)                                   \
  do {                                              \
    std::printf("[ RUN ] %s\n", #test);             \
    test();                                         \
  } while (0)
static int finish() {
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
