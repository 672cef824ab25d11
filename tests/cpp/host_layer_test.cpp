// host_layer_test.cpp -- exercises the C++ template layer (include/spsparse/*.hpp) the way user code of the
// reference would: same includes, same calls.  Needs a GPU.  Built by __graft_entry__.build(), run by
// tests/test_gpu_dropin.py.  Expected values come from the reference's own tests where cited, or from
// hand-checkable small cases.
#include <spsparse/VectorCooArray.hpp>
#include <spsparse/multiply_sparse.hpp>
#include <spsparse_b200/device_array.hpp>

#include <cmath>
#include <cstdio>

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

static std::vector<int> col(Mat const &m, int d) { return m.index_data(d); }

static void test_consolidate() {
    // tests/test_array.cpp:135-168
    Mat a({2, 4});
    a.add({1, 3}, 5.); a.add({1, 2}, 3.); a.add({0, 3}, 17.); a.add({0, 1}, 14.); a.add({1, 2}, 15.);
    Mat r(a.shape);
    consolidate(r, a, {0, 1});
    CHECK(same(col(r, 0), {0, 0, 1, 1}) && same(col(r, 1), {1, 3, 2, 3}) && same(r.val_data(), {14., 17., 18., 5.}));
    CHECK(!r.edit_mode && r.sort_order[0] == 0 && r.sort_order[1] == 1);
    CHECK(same(dim_beginnings(r), {(size_t)0, (size_t)2, (size_t)4}));
    r.clear();
    consolidate(r, a, {1, 0});
    CHECK(same(col(r, 0), {0, 1, 0, 1}) && same(col(r, 1), {1, 2, 3, 3}) && same(r.val_data(), {14., 18., 17., 5.}));
    CHECK(same(r.dim_beginnings(), {(size_t)0, (size_t)1, (size_t)2, (size_t)4}));
    // policies: first / last of the duplicates
    Mat f(a.shape), l(a.shape);
    consolidate(f, a, {0, 1}, DuplicatePolicy::LEAVE_ALONE);
    consolidate(l, a, {0, 1}, DuplicatePolicy::REPLACE);
    CHECK(f.val(2) == 3. && l.val(2) == 15.);
    // in-place form, no-op when already consolidated that way (VectorCooArray.hpp:306)
    a.consolidate({0, 1});
    CHECK(a.size() == 4 && !a.edit_mode);
    a.consolidate({0, 1});
    CHECK(a.size() == 4);
    // consolidate appends to `ret` (it does not clear it)
    Mat acc(a.shape);
    acc.add({1, 1}, 99.);
    consolidate(acc, a, {0, 1});
    CHECK(acc.size() == 5 && acc.val(0) == 99.);
    // zeros dropped on input, cancelling duplicates stay as an explicit 0
    Mat z({3, 3});
    z.add({1, 1}, 1.); z.add({1, 1}, -1.); z.add({2, 2}, 0.);
    z.consolidate({0, 1});
    CHECK(z.size() == 1 && z.val(0) == 0.);
    // zero_nan drops NaN in the leading run only (quirk 2)
    Mat nn({4, 4});
    nn.add({0, 0}, NAN); nn.add({1, 1}, 1.); nn.add({2, 2}, NAN);
    nn.consolidate({0, 1}, DuplicatePolicy::ADD, true);
    CHECK(nn.size() == 2 && nn.index(0, 0) == 1 && std::isnan(nn.val(1)));
    // empty input
    Mat e({4, 4}), er({4, 4});
    consolidate(er, e, {0, 1});
    CHECK(er.size() == 0 && !er.edit_mode);
    CHECK(dim_beginnings(er).empty());
}

static void test_permutation_and_rows() {
    // tests/test_array.cpp:67-79
    Mat a({2, 4});
    a.add({1, 3}, 5.); a.add({1, 2}, 3.); a.add({0, 3}, 17.);
    CHECK(same(sorted_permutation(a, {0, 1}), {(size_t)2, (size_t)1, (size_t)0}));
    CHECK(same(sorted_permutation(a, {1, 0}), {(size_t)1, (size_t)2, (size_t)0}));
    // tests/test_array.cpp:170-218
    Mat b({20, 10});
    b.add({1, 0}, 15.); b.add({1, 3}, 17.); b.add({2, 4}, 17.); b.add({6, 4}, 10.);
    b.consolidate({0, 1});
    std::vector<int> rows, firstcols;
    for (auto ii(b.dim_beginnings_xiter()); !ii.eof(); ++ii) {
        rows.push_back(*ii);
        auto jj(ii.sub_xiter());
        firstcols.push_back(*jj);
    }
    CHECK(same(rows, {1, 2, 6}) && same(firstcols, {0, 4, 4}));
}

static void test_errors() {
    Vec v({4});
    bool threw = false;
    try { v.add({17}, 4.); } catch (spsparse::Exception const &) { threw = true; }  // tests/test_array.cpp:49-56
    CHECK(threw);
    Mat u({3, 3});
    u.add({1, 1}, 1.);
    threw = false;
    try { dim_beginnings(u); } catch (spsparse::Exception const &) { threw = true; }  // algorithm.hpp:82-84
    CHECK(threw);
    Mat a({2, 3}), b({2, 2}), c;
    a.add({0, 0}, 1.); b.add({0, 0}, 1.);
    threw = false;
    try { multiply(c, 1.0, (Vec *)0, a, '.', (Vec *)0, b, '.', (Vec *)0); } catch (spsparse::Exception const &) { threw = true; }
    CHECK(threw);                                  // multiply_sparse.hpp:172-174
    CHECK(c.shape[0] == 2 && c.shape[1] == 2);     // shape is set before the check (:169)
}

static void test_multiply() {
    // tests/test_multiply_sparse.cpp:41-79 (disabled in the reference; answer verified on it)
    Mat row({2, 10});
    row.add({0, 8}, 6.); row.add({0, 4}, 4.); row.add({0, 0}, 2.); row.add({0, 3}, 3.); row.add({1, 8}, 3.);
    Vec scale({10});
    scale.add({0}, 2.); scale.add({4}, 4.); scale.add({8}, 4.);
    Mat colm({10, 1});
    colm.add({0, 0}, 2.); colm.add({3, 0}, 3.); colm.add({8, 0}, 5.);
    Vec eye({10});
    for (int i = 0; i < 10; ++i) eye.add({i}, 1.);
    Mat r;
    multiply(r, 1.0, &eye, row, '.', &scale, colm, '.', &eye);
    CHECK(r.size() == 2 && same(col(r, 0), {0, 1}) && same(col(r, 1), {0, 0}) && same(r.val_data(), {128., 60.}));
    CHECK(r.edit_mode && r.sort_order[0] == -1);   // left in edit mode, like the reference
    // transposes: (A B)^T == B^T A^T
    Mat A({3, 4}), B({4, 2});
    A.add({0, 1}, 2.); A.add({2, 3}, 3.); A.add({1, 1}, -1.); A.add({0, 3}, 5.);
    B.add({1, 0}, 7.); B.add({3, 1}, 11.); B.add({3, 0}, 1.);
    Mat AB, BtAt;
    multiply(AB, 1.0, (Vec *)0, A, '.', (Vec *)0, B, '.', (Vec *)0);
    multiply(BtAt, 1.0, (Vec *)0, B, 'T', (Vec *)0, A, 'T', (Vec *)0);
    CHECK(same(col(AB, 0), {0, 0, 1, 2, 2}) && same(col(AB, 1), {0, 1, 0, 0, 1}));
    CHECK(same(AB.val_data(), {19., 55., -7., 3., 33.}));
    BtAt.consolidate({1, 0});
    CHECK(same(col(BtAt, 1), {0, 0, 1, 2, 2}) && same(col(BtAt, 0), {0, 1, 0, 0, 1}) && same(BtAt.val_data(), {19., 55., -7., 3., 33.}));
    // exact-zero dot products are dropped, scale C applies
    Mat P({1, 2}), Q({2, 2}), PQ;
    P.add({0, 0}, 1.); P.add({0, 1}, 1.);
    Q.add({0, 0}, 1.); Q.add({1, 0}, -1.); Q.add({1, 1}, 4.);
    multiply(PQ, 2.5, (Vec *)0, P, '.', (Vec *)0, Q, '.', (Vec *)0);
    CHECK(PQ.size() == 1 && PQ.index(1, 0) == 1 && PQ.val(0) == 10.);
    // matrix * vector
    Vec V({4}), y;
    V.add({1}, 2.); V.add({3}, 1.); V.add({1}, 1.);
    multiply(y, 1.0, (Vec *)0, A, '.', (Vec *)0, V);
    CHECK(y.size() == 3 && same(y.index_data(0), {0, 1, 2}) && same(y.val_data(), {11., -3., 3.}));
    // generic accumulator: results delivered through add()
    ScalarAccumulator<Mat> total;
    Mat Acons(A.shape);
    consolidate(Acons, A, {0, 1});
    copy(total, Acons);
    CHECK(total.val == 9.);
}

static void test_mult_xiters() {
    // multiply_sparse.hpp:39-111: rows of a consolidated matrix, alone and joined with a scale vector
    Mat b({20, 10});
    b.add({6, 4}, 10.); b.add({1, 3}, 17.); b.add({2, 4}, 17.); b.add({1, 0}, 15.); b.add({9, 9}, 1.);
    b.consolidate({0, 1});
    std::vector<int> rows, cols;
    std::vector<double> sv, vals;
    auto plain(new_mult_xiter(b, (Vec *)0));
    for (; !plain->eof(); ++*plain) {
        rows.push_back(plain->index());
        sv.push_back(plain->scale_val());
        for (auto jj(plain->sub_xiter()); !jj.eof(); ++jj) { cols.push_back(*jj); vals.push_back(jj.val()); }
    }
    CHECK(same(rows, {1, 2, 6, 9}) && same(sv, {1., 1., 1., 1.}));
    CHECK(same(cols, {0, 3, 4, 4, 9}) && same(vals, {15., 17., 17., 10., 1.}));
    Vec s({20});
    s.add({0}, 7.); s.add({2}, 3.); s.add({5}, 4.); s.add({9}, .5); s.add({19}, 2.);
    rows.clear(); sv.clear(); cols.clear();
    auto scaled(new_mult_xiter(b, &s));
    for (; !scaled->eof(); ++*scaled) {
        rows.push_back(scaled->index());
        sv.push_back(scaled->scale_val());
        auto jj(scaled->sub_xiter());
        cols.push_back(*jj);
    }
    CHECK(same(rows, {2, 9}) && same(sv, {3., .5}) && same(cols, {4, 9}));
    // column-major operand: the walker follows the leading sorted dimension (columns), entries report rows
    b.consolidate({1, 0});
    rows.clear(); cols.clear();
    SimpleMultXiter<Mat> bycol(b);
    for (; !bycol.eof(); ++bycol) {
        cols.push_back(bycol.index());
        for (auto jj(bycol.sub_xiter()); !jj.eof(); ++jj) rows.push_back(*jj);
    }
    CHECK(same(cols, {0, 3, 4, 9}) && same(rows, {1, 1, 2, 6, 9}));
}

static void test_device_resident_chain() {
    // b200::DeviceCooArray: the same operations with the array kept on the GPU between them; every step is compared with the
    // host-container call of the same name
    typedef b200::DeviceCooArray<2> DMat;
    typedef b200::DeviceCooArray<1> DVec;
    Mat a({3, 4});
    a.add({2, 1}, 4.); a.add({0, 3}, 1.5); a.add({2, 1}, -1.); a.add({1, 0}, 0.); a.add({0, 0}, 2.); a.add({2, 3}, 7.);
    DMat da(a);
    CHECK(da.size() == 6 && da.shape()[0] == 3 && da.shape()[1] == 4 && da.sort_order()[0] == -1);
    // copy / transpose (algorithm.hpp:30-57): entries keep their order
    {
        Mat want({4, 3}), got({4, 3});
        transpose(want, a, {1, 0});
        da.transpose({1, 0}).download(got);
        CHECK(col(got, 0) == col(want, 0) && col(got, 1) == col(want, 1) && got.val_data() == want.val_data());
        Mat cp({3, 4});
        da.copy().download(cp);
        CHECK(col(cp, 0) == col(a, 0) && col(cp, 1) == col(a, 1) && cp.val_data() == a.val_data());
    }
    // consolidate on the device == consolidate through the host container; the flag travels
    DMat dc = da.consolidate(ROW_MAJOR);
    {
        Mat want(a.shape), got(a.shape);
        consolidate(want, a, ROW_MAJOR);
        dc.download(got);
        CHECK(dc.sort_order()[0] == 0 && dc.sort_order()[1] == 1);
        CHECK(col(got, 0) == col(want, 0) && col(got, 1) == col(want, 1) && got.val_data() == want.val_data());
        CHECK(same(got.val_data(), {2., 1.5, 3., 7.}));                       // the zero dropped, the duplicate summed
        CHECK(same(dc.dim_beginnings(), {size_t(0), size_t(2), size_t(4)}));  // rows 0 and 2, and the sentinel
    }
    // to_dense with the three duplicate policies of DenseAccum (accum.hpp:110-140), from_dense = to_sparse
    {
        std::vector<double> add = da.to_dense(), repl = da.to_dense(DuplicatePolicy::REPLACE);
        CHECK(add.size() == 12 && add[2 * 4 + 1] == 3. && repl[2 * 4 + 1] == -1. && add[0] == 2. && add[3] == 1.5 && add[11] == 7. && add[4] == 0.);
        DMat back = DMat::from_dense({3, 4}, add.data());
        Mat got({3, 4});
        back.download(got);
        CHECK(same(col(got, 0), {0, 0, 2, 2}) && same(col(got, 1), {0, 3, 1, 3}) && same(got.val_data(), {2., 1.5, 3., 7.}));
    }
    // multiply, matrix x matrix and matrix x vector, against the host-container calls
    {
        Mat b({4, 2});
        b.add({3, 1}, 2.); b.add({0, 0}, 3.); b.add({1, 1}, 5.); b.add({3, 0}, -1.);
        Vec w({4}), v({4});
        w.add({0}, 2.); w.add({1}, 1.); w.add({3}, 0.5);
        v.add({1}, 10.); v.add({3}, 1.);
        DMat db(b);
        DVec dw(w), dv(v);
        Mat want, got;
        multiply(want, 2., (Vec *)0, a, '.', &w, b, '.', (Vec *)0);
        got.set_shape(want.shape);
        b200::multiply(2., nullptr, da, '.', &dw, db, '.', nullptr).download(got);
        CHECK(got.size() == want.size() && col(got, 0) == col(want, 0) && col(got, 1) == col(want, 1) && got.val_data() == want.val_data());
        Vec wantv, gotv;
        multiply(wantv, 1., (Vec *)0, a, '.', (Vec *)0, v);
        gotv.set_shape(wantv.shape);
        b200::multiply(1., nullptr, da, '.', nullptr, dv).download(gotv);
        CHECK(gotv.size() == wantv.size() && gotv.index_data(0) == wantv.index_data(0) && gotv.val_data() == wantv.val_data());
        // A^T * A through a device transpose + consolidate equals the 'T' flag
        Mat t1, t2;
        multiply(t1, 1., (Vec *)0, a, 'T', (Vec *)0, a, '.', (Vec *)0);
        t2.set_shape(t1.shape);
        b200::multiply(1., nullptr, da.transpose({1, 0}).consolidate(ROW_MAJOR), '.', nullptr, dc, '.', nullptr).download(t2);
        CHECK(t2.size() == t1.size() && col(t2, 0) == col(t1, 0) && col(t2, 1) == col(t1, 1) && t2.val_data() == t1.val_data());
    }
    // moves, empties, errors
    {
        DMat moved(std::move(dc));
        CHECK(dc.handle() == nullptr && moved.size() == 4);
        DMat empty;
        CHECK(empty.size() == 0 && empty.dim_beginnings().empty());
        bool threw = false;
        try { (void)da.dim_beginnings(); } catch (spsparse::Exception const &) { threw = true; }   // not flagged sorted
        CHECK(threw);
    }
}

int main() {
    test_device_resident_chain();
    test_mult_xiters();
    test_consolidate();
    test_permutation_and_rows();
    test_errors();
    test_multiply();
    std::printf("host_layer_test: %d failure(s)\n", failures);
    return failures ? 1 : 0;
}
