// Minimal stand-in for <gtest/gtest.h> (gtest is absent from this image).  TEST INFRASTRUCTURE
// ONLY.  Just enough for the reference's own test sources (tests/test_xiter.cpp,
// tests/test_array.cpp, tests/test_multiply_sparse.cpp) to compile unmodified.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

class Test {
public:
    virtual ~Test() {}
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};

struct Registry {
    struct Entry { std::string name; std::function<Test *()> make; };
    static std::vector<Entry> &tests() { static std::vector<Entry> t; return t; }
    static int &failures() { static int f = 0; return f; }
};

struct Registrar {
    Registrar(const char *suite, const char *name, std::function<Test *()> make) {
        Registry::tests().push_back({std::string(suite) + "." + name, make});
    }
};

// Swallows `<< ...` after a failed expectation / FAIL().
struct Message {
    bool active;
    explicit Message(bool a) : active(a) {}
    template <class T> Message &operator<<(T const &v) { if (active) std::cerr << v; return *this; }
    Message &operator<<(std::ostream &(*m)(std::ostream &)) { if (active) std::cerr << m; return *this; }
    ~Message() { if (active) std::cerr << std::endl; }
};

inline Message fail_at(const char *file, int line, const char *what) {
    ++Registry::failures();
    std::cerr << file << ":" << line << ": Failure: " << what << " ";
    return Message(true);
}

// gtest's AlmostEquals: within 4 ULPs (and never equal if either is NaN)
inline bool double_eq_4ulp(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return false;
    auto biased = [](double d) {
        uint64_t u; std::memcpy(&u, &d, 8);
        const uint64_t sign = 1ull << 63;
        return (u & sign) ? (~u + 1) : (u | sign);
    };
    uint64_t x = biased(a), y = biased(b);
    return (x > y ? x - y : y - x) <= 4;
}

// `return Voidify() = Message << ...;` is legal in a void function (gtest's own trick).
struct Voidify { void operator=(Message const &) {} };

inline void InitGoogleTest(int *, char **) {}

inline int RunAll() {
    int nfail_tests = 0;
    for (auto &e : Registry::tests()) {
        int before = Registry::failures();
        Test *t = e.make();
        t->SetUp(); t->TestBody(); t->TearDown();
        delete t;
        bool ok = Registry::failures() == before;
        std::printf("[%s] %s\n", ok ? "  OK  " : "FAILED", e.name.c_str());
        if (!ok) ++nfail_tests;
    }
    std::printf("%d test(s), %d failed\n", (int)Registry::tests().size(), nfail_tests);
    return nfail_tests ? 1 : 0;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAll()

#define TEST_F(fixture, name)                                                               \
    class fixture##_##name##_Test : public fixture { public: void TestBody() override; };   \
    static ::testing::Registrar fixture##_##name##_reg(                                     \
        #fixture, #name, []() -> ::testing::Test * { return new fixture##_##name##_Test; }); \
    void fixture##_##name##_Test::TestBody()

#define GT_CHECK_(cond, text) \
    if (cond) ; else ::testing::fail_at(__FILE__, __LINE__, text)

#define EXPECT_EQ(a, b) GT_CHECK_((a) == (b), "EXPECT_EQ(" #a ", " #b ")")
#define EXPECT_NE(a, b) GT_CHECK_((a) != (b), "EXPECT_NE(" #a ", " #b ")")
#define EXPECT_TRUE(a) GT_CHECK_((a), "EXPECT_TRUE(" #a ")")
#define EXPECT_DOUBLE_EQ(a, b) GT_CHECK_(::testing::double_eq_4ulp((a), (b)), "EXPECT_DOUBLE_EQ(" #a ", " #b ")")
#define FAIL() return ::testing::Voidify() = ::testing::fail_at(__FILE__, __LINE__, "FAIL()")
