// host_only_test.cpp -- the parts of the C++ template layer (include/spsparse/*.hpp) that never reach the GPU:
// the container and its accessors/iterators, the sorted joins (a8), the predicates and
// constants (a6), the accumulators and the error convention (a12).  Runs in the CPU suite
// (tests/test_host_layer_cpu.py); everything that computes on the device is in host_layer_test.cpp.
// Expected values: the reference's own tests where cited, hand-checkable otherwise.
#include <spsparse/VectorCooArray.hpp>
#include <spsparse/multiply_sparse.hpp>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <sstream>

using namespace spsparse;

static int failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) { ++failures; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

typedef VectorCooArray<int, double, 2> Mat;
typedef VectorCooArray<int, double, 1> Vec;

template <class T>
static bool same(std::vector<T> const &a, std::initializer_list<T> b) { return a == std::vector<T>(b); }

static Mat sample() {
    Mat a({4, 6});
    a.add({1, 3}, 5.); a.add({1, 2}, 3.); a.add({0, 3}, 17.); a.add({3, 5}, -2.);
    return a;
}

static void test_container() {
    Mat a(sample());
    CHECK(Mat::rank == 2 && a.shape[0] == 4 && a.shape[1] == 6 && a.size() == 4);
    CHECK(a.edit_mode && a.sort_order[0] == -1);                       // VectorCooArray.hpp:29-34
    CHECK(a.index(0, 2) == 0 && a.index(1, 2) == 3 && a.val(2) == 17.);
    CHECK((a.index(3) == std::array<int, 2>{3, 5}));
    CHECK(same(a.index_vec(1), {1, 2}));
    a.set_index(3, {2, 4});
    a.val(3) = 8.;
    a.index(1, 0) = 1;
    CHECK(same(a.index_data(0), {1, 1, 0, 2}) && same(a.index_data(1), {1, 2, 3, 4}) && same(a.val_data(), {5., 3., 17., 8.}));
    a.reserve(100);
    CHECK(a.size() == 4);
    // set_sorted leaves edit mode, edit() re-enters it and forgets the order (:131-135, :97-101)
    a.set_sorted({0, 1});
    CHECK(!a.edit_mode && a.sort_order[0] == 0 && a.sort_order[1] == 1);
    bool threw = false;
    try { a.add({0, 0}, 1.); } catch (Exception const &) { threw = true; }   // :241-243
    CHECK(threw && a.size() == 4);
    a.edit();
    CHECK(a.edit_mode && a.sort_order[0] == -1);
    a.add({0, 0}, 1.);
    CHECK(a.size() == 5);
    // blank arrays of the same shape
    Mat b(a.make_blank());
    std::unique_ptr<Mat> c(a.new_blank());
    CHECK(b.size() == 0 && b.shape == a.shape && c->size() == 0 && c->shape == a.shape);
    b.set_shape({7, 8});
    CHECK(b.shape[0] == 7 && b.shape[1] == 8);
    // copies are deep, moves carry the storage
    Mat d(a);
    d.val(0) = 99.;
    CHECK(a.val(0) == 5.);
    Mat e;
    e = std::move(d);
    CHECK(e.size() == 5 && e.val(0) == 99.);
    e = a;
    CHECK(e.val(0) == 5.);
    a.clear();
    CHECK(a.size() == 0 && a.edit_mode && e.size() == 5);
    // aliases (:352-356)
    VectorCooMatrix<int, double> m2({2, 2});
    VectorCooVector<int, double> v1({3});
    CHECK(m2.rank == 2 && v1.rank == 1);
}

static void test_bounds() {
    // tests/test_array.cpp:49-56
    Vec v({4});
    v.add({1}, 2.);
    bool threw = false;
    try { v.add({17}, 4.); } catch (Exception const &) { threw = true; }
    CHECK(threw);
    threw = false;
    try { v.add({-1}, 4.); } catch (Exception const &) { threw = true; }
    CHECK(threw && v.size() == 1);
    Mat m({2, 3});
    threw = false;
    try { m.add({1, 3}, 1.); } catch (Exception const &) { threw = true; }
    CHECK(threw);
    threw = false;
    try { m.add({2, 0}, 1.); } catch (Exception const &) { threw = true; }
    CHECK(threw && m.size() == 0);
    m.add({1, 2}, 1.);
    CHECK(m.size() == 1);
}

static void test_iterators() {
    Mat a(sample());
    // tests/test_array.cpp:81-107: *it is the index tuple, it.val() the value, it.index(k) one coordinate
    std::vector<int> rows, cols;
    std::vector<double> vals;
    for (auto ii(a.begin()); ii != a.end(); ++ii) {
        rows.push_back((*ii)[0]);
        cols.push_back(ii.index(1));
        vals.push_back(ii.val());
    }
    CHECK(same(rows, {1, 1, 0, 3}) && same(cols, {3, 2, 3, 5}) && same(vals, {5., 3., 17., -2.}));
    auto it(a.begin());
    CHECK(it.offset() == 0 && (it[2] == std::array<int, 2>{0, 3}));
    it += 3;
    CHECK(it.offset() == 3 && it.val() == -2.);
    --it;
    it -= 1;
    CHECK(it.offset() == 1 && (it + 2).offset() == 3 && it.index() == a.index(1));
    it.val() = 4.;
    it.set_index({2, 2});
    CHECK(a.val(1) == 4. && a.index(0, 1) == 2 && a.index(1, 1) == 2);
    Mat const &ca(a);
    int count = 0;
    for (auto ii(ca.cbegin()); ii != ca.cend(); ++ii) ++count;
    CHECK(count == 4);
    CHECK(ca.begin(1).offset() == 1 && a.begin(2).offset() == 2 && a.end().offset() == 4);
    // one dimension at a time (array.hpp:47-67)
    std::vector<int> d1;
    std::vector<double> dv;
    for (auto ii(ca.dim_begin(1)); ii != ca.dim_end(1); ++ii) { d1.push_back(*ii); dv.push_back(ii.val()); }
    CHECK(same(d1, {3, 2, 3, 5}) && same(dv, {5., 4., 17., -2.}));
    CHECK(*ca.dim_iter(0, 2) == 0);
    // printing
    std::ostringstream os;
    Vec v({3});
    v.add({2}, 1.5);
    os << v;
    CHECK(os.str() == "VectorCooArray<{3}>((2 : 1.5))");
}

static void test_joins() {
    typedef STLXiter<std::vector<int>::iterator> X;
    // tests/test_xiter.cpp:52-125
    std::vector<int> v1 = {0, 2, 4, 6}, v2 = {0, 1, 2, 3, 4, 5, 6, 7}, v3 = {1, 2, 3, 6}, out;
    for (auto ii(join2_xiter(X(v1.begin(), v1.end()), X(v2.begin(), v2.end()))); !ii.eof(); ++ii) {
        CHECK(*ii.i1 == *ii.i2);
        out.push_back(*ii.i1);
    }
    CHECK(same(out, {0, 2, 4, 6}));
    out.clear();
    for (auto ii(join2_xiter(X(v2.begin(), v2.end()), X(v1.begin(), v1.end()))); !ii.eof(); ++ii) out.push_back(*ii.i2);
    CHECK(same(out, {0, 2, 4, 6}));
    std::vector<int> w1 = {0, 2, 4, 5, 6, 7, 8, 9}, w2 = {1, 2, 3, 4, 6};
    out.clear();
    for (auto ii(join2_xiter(X(w1.begin(), w1.end()), X(w2.begin(), w2.end()))); !ii.eof(); ++ii) out.push_back(*ii.i1);
    CHECK(same(out, {2, 4, 6}));
    out.clear();
    for (auto ii(join3_xiter(X(v1.begin(), v1.end()), X(v2.begin(), v2.end()), X(v3.begin(), v3.end()))); !ii.eof(); ++ii) {
        CHECK(*ii.i1 == *ii.i2 && *ii.i1 == *ii.i3);
        out.push_back(*ii.i3);
    }
    CHECK(same(out, {2, 6}));
    // no common element / an empty input / a single element
    std::vector<int> odd = {1, 3, 5}, even = {0, 2, 4}, none, one = {4};
    CHECK(join2_xiter(X(odd.begin(), odd.end()), X(even.begin(), even.end())).eof());
    CHECK(join2_xiter(X(none.begin(), none.end()), X(even.begin(), even.end())).eof());
    CHECK(join2_xiter(X(even.begin(), even.end()), X(none.begin(), none.end())).eof());
    CHECK(join3_xiter(X(even.begin(), even.end()), X(v2.begin(), v2.end()), X(none.begin(), none.end())).eof());
    auto j(join3_xiter(X(one.begin(), one.end()), X(even.begin(), even.end()), X(v2.begin(), v2.end())));
    CHECK(!j.eof() && *j.i1 == 4 && j.i2.offset() == 2 && j.i3.offset() == 4 && j.total_in_use == 3);
    ++j;
    CHECK(j.eof());
    // xiters over a sparse vector expose the matching value: the dot product of two sparse vectors
    Vec a({10}), b({10});
    a.add({1}, 2.); a.add({4}, 3.); a.add({7}, 5.);
    b.add({0}, 1.); b.add({4}, 10.); b.add({7}, 100.); b.add({9}, 7.);
    double dot = 0;
    for (auto ii(join2_xiter(make_val_xiter(a.dim_begin(0), a.dim_end(0)), make_val_xiter(b.dim_begin(0), b.dim_end(0))));
         !ii.eof(); ++ii)
        dot += ii.i1.val() * ii.i2.val();
    CHECK(dot == 530.);
    X x(v1.begin(), v1.end());
    ++x; ++x;
    CHECK(x.offset() == 2 && *x == 4 && !x.eof());
}

static void test_predicates_and_constants() {
    CHECK(isnone(0.) && isnone(-0.) && !isnone(1e-300) && !isnone(NAN) && isnone(NAN, true) && !isnone(INFINITY, true));
    CHECK(isnone(0) && !isnone(3));                                               // spsparse.hpp:95-103
    CHECK(ROW_MAJOR[0] == 0 && ROW_MAJOR[1] == 1 && COL_MAJOR[0] == 1 && COL_MAJOR[1] == 0);   // spsparse.cpp:30-31
    CHECK((int)DuplicatePolicy::LEAVE_ALONE == 0 && (int)DuplicatePolicy::ADD == 1 && (int)DuplicatePolicy::REPLACE == 2);
    CHECK(b200::policy_code(DuplicatePolicy::LEAVE_ALONE) == SPB_LEAVE_ALONE && b200::policy_code(DuplicatePolicy::ADD) == SPB_ADD &&
          b200::policy_code(DuplicatePolicy::REPLACE) == SPB_REPLACE);
}

struct Recorder {  // a rank-1 accumulator that logs what reaches it
    static const int rank = 1;
    typedef int index_type;
    typedef double val_type;
    std::vector<int> *log;
    double *sum;
    void add(std::array<int, 1> const &ix, double const &v) { log->push_back(ix[0]); *sum += v; }
};

static void test_accumulators_and_host_loops() {
    Mat a(sample());
    // copy / transpose go entry by entry through add(), in storage order (algorithm.hpp:30-57)
    Mat c(a.shape), t({6, 4});
    copy(c, a);
    CHECK(c.index_data(0) == a.index_data(0) && c.index_data(1) == a.index_data(1) && c.val_data() == a.val_data());
    transpose(t, a, {1, 0});
    CHECK(t.index_data(0) == a.index_data(1) && t.index_data(1) == a.index_data(0) && t.val_data() == a.val_data());
    ScalarAccumulator<Mat> total;
    copy(total, a);
    CHECK(total.val == 23.);
    // in-place transpose overwrites the entries (VectorCooArray.hpp:144-148); a square shape keeps them in bounds
    Mat s({6, 6});
    s.add({1, 3}, 5.); s.add({0, 2}, 7.);
    s.transpose({1, 0});
    CHECK(same(s.index_data(0), {3, 2}) && same(s.index_data(1), {1, 0}) && same(s.val_data(), {5., 7.}));
    // PermuteAccum: pick / reorder dimensions on the way into another accumulator (accum.hpp:73-93)
    std::vector<int> seen;
    double seen_sum = 0;
    PermuteAccum<2, Recorder> pick(Recorder{&seen, &seen_sum}, {1});
    copy(pick, a);
    CHECK(same(seen, {3, 2, 3, 5}) && seen_sum == 23.);
    // Consolidate<> leaves an array alone that is already flagged sorted that way (algorithm.hpp:354-369)
    Mat sorted({3, 3});
    sorted.add({0, 1}, 1.); sorted.add({2, 0}, 2.);
    sorted.set_sorted({0, 1});
    Consolidate<Mat> keep(&sorted, {0, 1});
    CHECK(&keep() == &sorted);
    // the in-place form is a no-op in the same situation (VectorCooArray.hpp:306)
    sorted.consolidate({0, 1});
    CHECK(sorted.size() == 2 && !sorted.edit_mode);
}

// A consolidated matrix whose row-start list is supplied by the test instead of by the device (the list is a lazily
// filled cache, VectorCooArray.hpp:323-335): lets the row walkers run without a GPU.
struct PresetMat : Mat {
    PresetMat(std::array<size_t, 2> const &shape) : Mat(shape) {}
    void preset(std::array<int, 2> const &order, std::vector<size_t> const &starts) {
        set_sorted(order);
        _dim_beginnings = starts;
        dim_beginnings_set = true;
    }
};

static void test_mult_xiters() {
    // multiply_sparse.hpp:39-111
    PresetMat b({20, 10});
    b.add({1, 0}, 15.); b.add({1, 3}, 17.); b.add({2, 4}, 17.); b.add({6, 4}, 10.); b.add({9, 9}, 1.);
    b.preset({0, 1}, {0, 2, 3, 4, 5});
    Mat const &bm(b);
    std::vector<int> rows, cols;
    std::vector<double> sv, vals;
    auto plain(new_mult_xiter(bm, (Vec *)0));
    for (; !plain->eof(); ++*plain) {
        rows.push_back(plain->index());
        sv.push_back(plain->scale_val());
        for (auto jj(plain->sub_xiter()); !jj.eof(); ++jj) { cols.push_back(*jj); vals.push_back(jj.val()); }
    }
    CHECK(same(rows, {1, 2, 6, 9}) && same(sv, {1., 1., 1., 1.}));
    CHECK(same(cols, {0, 3, 4, 4, 9}) && same(vals, {15., 17., 17., 10., 1.}));
    // with a scale vector: only rows the vector has an entry for, each with its scale value
    Vec s({20});
    s.add({0}, 7.); s.add({2}, 3.); s.add({5}, 4.); s.add({9}, .5); s.add({19}, 2.);
    rows.clear(); sv.clear(); cols.clear();
    auto scaled(new_mult_xiter(bm, &s));
    for (; !scaled->eof(); ++*scaled) {
        rows.push_back(scaled->index());
        sv.push_back(scaled->scale_val());
        auto jj(scaled->sub_xiter());
        cols.push_back(*jj);
    }
    CHECK(same(rows, {2, 9}) && same(sv, {3., .5}) && same(cols, {4, 9}));
    Vec miss({20});
    miss.add({0}, 1.); miss.add({3}, 1.); miss.add({19}, 1.);
    ScaledMultXiter<Mat, Vec> none(bm, miss);
    CHECK(none.eof());
    // a column-major operand is walked by column; the entries of a column report their rows
    PresetMat c({20, 10});
    c.add({1, 0}, 15.); c.add({1, 3}, 17.); c.add({2, 4}, 17.); c.add({6, 4}, 10.); c.add({9, 9}, 1.);
    c.preset({1, 0}, {0, 1, 2, 4, 5});
    rows.clear(); cols.clear();
    SimpleMultXiter<Mat> bycol(c);
    for (; !bycol.eof(); ++bycol) {
        cols.push_back(bycol.index());
        for (auto jj(bycol.sub_xiter()); !jj.eof(); ++jj) rows.push_back(*jj);
    }
    CHECK(same(cols, {0, 3, 4, 9}) && same(rows, {1, 1, 2, 6, 9}));
}

static int hook_calls = 0;
static char hook_text[256];
static void counting_hook(int, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(hook_text, sizeof hook_text, fmt, ap);
    va_end(ap);
    ++hook_calls;
    throw 42;
}

static void test_error_convention() {
    // the hook is a replaceable global that must not return (spsparse.hpp:47-54, spsparse.cpp:12-28)
    error_ptr saved = spsparse_error;
    spsparse_error = &counting_hook;
    Mat a({2, 3}), b({2, 2}), c, u({3, 3});
    a.add({0, 0}, 1.); b.add({0, 0}, 1.); u.add({1, 1}, 1.);
    int caught = 0;
    try { a.add({5, 5}, 1.); } catch (int) { ++caught; }
    CHECK(std::string(hook_text).find("out of bounds") != std::string::npos);
    try { dim_beginnings(u); } catch (int) { ++caught; }                              // algorithm.hpp:82-84: checked before any work
    CHECK(std::string(hook_text).find("sorted first") != std::string::npos);
    try { multiply(c, 1.0, (Vec *)0, a, '.', (Vec *)0, b, '.', (Vec *)0); } catch (int) { ++caught; }   // multiply_sparse.hpp:172-174
    CHECK(std::string(hook_text) == "Inner dimensions for A (3) and B (2) must match!");
    CHECK(c.shape[0] == 2 && c.shape[1] == 2);                                          // shape is set before the check (:169)
    Vec v({5}), y;
    v.add({0}, 1.);
    try { multiply(y, 1.0, (Vec *)0, a, '.', (Vec *)0, v); } catch (int) { ++caught; }  // :294-296
    CHECK(caught == 4 && hook_calls == 4 && y.shape[0] == 2);
    spsparse_error = saved;
    // empty short-circuits never reach the device (multiply_sparse.hpp:176-185): empty operand, empty scale, C == 0
    Mat e({3, 2}), r;
    Mat a2({2, 3});
    a2.add({0, 0}, 1.);
    Vec empty_scale({3});
    multiply(r, 1.0, (Vec *)0, a2, '.', (Vec *)0, e, '.', (Vec *)0);
    CHECK(r.size() == 0 && r.shape[0] == 2 && r.shape[1] == 2);
    e.add({0, 0}, 1.);
    multiply(r, 1.0, (Vec *)0, a2, '.', &empty_scale, e, '.', (Vec *)0);
    CHECK(r.size() == 0);
    multiply(r, 0.0, (Vec *)0, a2, '.', (Vec *)0, e, '.', (Vec *)0);
    CHECK(r.size() == 0);
    // consolidate of an empty array only flags the result (algorithm.hpp:263, 318)
    Mat z({4, 4}), zr({4, 4});
    consolidate(zr, z, {1, 0});
    CHECK(zr.size() == 0 && !zr.edit_mode && zr.sort_order[0] == 1 && zr.sort_order[1] == 0);
    CHECK(dim_beginnings(zr).empty());
    CHECK(sorted_permutation(z, {0, 1}).empty());
}

static void test_grow_raw_large() {
    // room for a multi-million-entry device result at the end of an array that already holds entries: the new pages are
    // pre-faulted in parallel (spb_host_prefault), the old entries stay, the new slots are value-initialised and writable
    Mat a({4, 6});
    a.add({1, 3}, 5.); a.add({0, 1}, -2.);
    const size_t n = (size_t(1) << 22) + 5;
    int *ip[2] = {nullptr, nullptr};
    double *vp = nullptr;
    a.grow_raw(n, ip, &vp);
    CHECK(a.size() == n + 2 && ip[0] == a.index_data(0).data() + 2 && ip[1] == a.index_data(1).data() + 2 && vp == a.val_data().data() + 2);
    CHECK(a.index(0, 0) == 1 && a.index(1, 1) == 1 && a.val(0) == 5. && a.val(1) == -2.);
    bool zero = true;
    for (size_t t = 0; t < n; t += 4099) zero = zero && ip[0][t] == 0 && ip[1][t] == 0 && vp[t] == 0.;
    CHECK(zero && ip[0][n - 1] == 0 && vp[n - 1] == 0.);
    vp[n - 1] = 3.5; ip[1][n - 1] = 5;
    CHECK(a.val(n + 1) == 3.5 && a.index(1, n + 1) == 5);
    // the helper itself: any range, any alignment, nothing but zeros written to the first byte of every page
    std::vector<char> buf(3 * 4096 + 17, 7);
    CHECK(spb_host_prefault(buf.data() + 5, buf.size() - 9) == 0 && spb_host_prefault(nullptr, 10) == 0 && spb_host_prefault(buf.data(), 0) == 0);
    CHECK(buf[5] == 0 && buf[4] == 7 && buf[6] == 7 && buf[buf.size() - 1] == 7);
}

int main() {
    test_grow_raw_large();
    test_container();
    test_bounds();
    test_iterators();
    test_joins();
    test_predicates_and_constants();
    test_accumulators_and_host_loops();
    test_mult_xiters();
    test_error_convention();
    std::printf("host_only_test: %d failure(s)\n", failures);
    return failures ? 1 : 0;
}
