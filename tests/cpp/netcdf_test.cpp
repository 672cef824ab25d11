// netcdf_test.cpp -- NetCDF I/O of COO arrays (reference slib/spsparse/netcdf.hpp, tests/test_netcdf.cpp)
// through include/spsparse/netcdf.hpp and the minimal classic-format netCDF of include/spsparse_b200/mini_netcdf/.
// Host only.  Sub-commands (tests/test_netcdf_cpu.py drives them):
//   selftest DIR        round trips, conversions, hyperslabs, error paths; prints "<n> failure(s)"
//   write PATH          a fixed VectorCooArray through ncio_spsparse (CDF-5, the reference's layout)
//   write-classic PATH FORMAT   generic variables with classic types (FORMAT: classic | classic64) for an independent reader
//   dump PATH           canonical text of any classic-format file (written by an independent writer)
#include <spsparse/VectorCooArray.hpp>
#include <spsparse/netcdf.hpp>

#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <string>

using namespace spsparse;
using namespace netCDF;

static int failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) { ++failures; std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

typedef VectorCooArray<int, double, 2> Mat;
typedef VectorCooArray<int, double, 1> Vec;

static Mat fixed_array() {
    Mat a({5, 6});   // the array of the reference's tests/test_netcdf.cpp:51-54, plus a duplicate, a zero and a negative value
    a.add({1, 2}, 2.);
    a.add({3, 3}, 6.);
    a.add({4, 5}, 1.);
    a.add({1, 2}, -0.25);
    a.add({0, 0}, 0.);
    return a;
}

template <class A>
static bool same_entries(A const &x, A const &y) {
    if (x.size() != y.size()) return false;
    auto i = x.begin();
    auto j = y.begin();
    for (; i != x.end(); ++i, ++j)
        if (i.index() != j.index() || std::memcmp(&i.val(), &j.val(), sizeof(double)) != 0) return false;
    return true;
}

static void test_round_trip(std::string const &dir) {
    // tests/test_netcdf.cpp:49-98: write, read with alloc, read without alloc
    const std::string fname = dir + "/round_trip.nc";
    Mat arr1(fixed_array());
    {
        ibmisc::NcIO ncio(fname, NcFile::replace);
        ncio_spsparse(ncio, arr1, true, "arr1");
        ncio.close();
    }
    Mat arr2;
    {
        ibmisc::NcIO ncio(fname, NcFile::read);
        ncio_spsparse(ncio, arr2, true, "arr1");
        ncio.close();
    }
    Mat arr3(arr1.shape);
    {
        ibmisc::NcIO ncio(fname, NcFile::read);
        ncio_spsparse(ncio, arr3, false, "arr1");
        ncio.close();
    }
    CHECK(arr2.shape == arr1.shape && same_entries(arr1, arr2) && same_entries(arr1, arr3));
    // the file as the reference lays it out (netcdf.hpp:93-106)
    {
        NcFile f(fname, NcFile::read);
        CHECK(f.getDim("arr1.size").getSize() == 5 && f.getDim("arr1.rank").getSize() == 2);
        NcVar info = f.getVar("arr1.info"), ind = f.getVar("arr1.indices"), vals = f.getVar("arr1.vals");
        CHECK(!info.isNull() && info.getType() == ncInt64 && info.getDimCount() == 0);
        CHECK(info.getAtt("shape").getType() == ncUint64 && info.getAtt("shape").getAttLength() == 2);
        unsigned long long shp[2] = {0, 0};
        info.getAtt("shape").getValues(shp);
        CHECK(shp[0] == 5 && shp[1] == 6);
        CHECK(ind.getType() == ncInt64 && ind.getDimCount() == 2 && ind.getDim(0).getName() == "arr1.size" && ind.getDim(1).getName() == "arr1.rank");
        CHECK(vals.getType() == ncDouble && vals.getDimCount() == 1);
        long long all[10];
        ind.getVar(all);
        const long long want[10] = {1, 2, 3, 3, 4, 5, 1, 2, 0, 0};
        CHECK(std::memcmp(all, want, sizeof want) == 0);
        // one entry at a time with conversion to int, as the reference's reader does (netcdf.hpp:70-75)
        int one[2];
        double v;
        ind.getVar({2, 0}, {1, 2}, one);
        vals.getVar({2, 0}, {1, 2}, &v);   // (start/count longer than the variable's rank, as the reference passes them)
        CHECK(one[0] == 4 && one[1] == 5 && v == 1.);
    }
    // two arrays in one file, the second of rank 1; an empty array
    const std::string f2 = dir + "/two.nc";
    Vec v1({7});
    v1.add({6}, 1.5);
    v1.add({0}, -3.);
    Mat empty({3, 3});
    {
        ibmisc::NcIO ncio(f2, NcFile::replace);
        ncio_spsparse(ncio, arr1, true, "A");
        ncio_spsparse(ncio, v1, true, "v");
        ncio.close();
    }
    {
        Mat a;
        Vec v;
        ibmisc::NcIO ncio(f2, NcFile::read);
        ncio_spsparse(ncio, a, true, "A");
        ncio_spsparse(ncio, v, true, "v");
        ncio.close();
        CHECK(same_entries(a, arr1) && v.shape == v1.shape && same_entries(v, v1));
    }
    // rank mismatch is reported through spsparse_error (netcdf.hpp:117-121)
    {
        bool threw = false;
        Vec wrong;
        try {
            ibmisc::NcIO ncio(f2, NcFile::read);
            ncio_spsparse(ncio, wrong, true, "A");
            ncio.close();
        } catch (spsparse::Exception const &) { threw = true; }
        CHECK(threw);
    }
    // more entries than one transfer block
    {
        Mat big({1000, 1000000});
        const size_t n = b200::NC_BLOCK * 2 + 12345;
        big.reserve(n);
        for (size_t t = 0; t < n; ++t) big.add({(int)(t % 1000), (int)((t * 7919) % 1000000)}, (double)t * 0.5 - 3.);
        const std::string f3 = dir + "/big.nc";
        {
            ibmisc::NcIO ncio(f3, NcFile::replace);
            ncio_spsparse(ncio, big, true, "M");
            ncio.close();
        }
        Mat back;
        ibmisc::NcIO ncio(f3, NcFile::read);
        ncio_spsparse(ncio, back, true, "M");
        ncio.close();
        CHECK(back.shape == big.shape && same_entries(back, big));
        std::remove(f3.c_str());
    }
}

static void test_mini_api(std::string const &dir) {
    const std::string fname = dir + "/api.nc";
    {
        NcFile f(fname, NcFile::replace, NcFile::classic64);
        NcDim dy = f.addDim("y", 3), dx = f.addDim("x", 4);
        NcVar g = f.addVar("grid", ncInt, {dy, dx});
        NcVar s = f.addVar("scale", ncFloat, dx);
        NcVar b = f.addVar("bytes", ncByte, dy);
        f.putAtt("title", std::string("mini"));
        g.putAtt("units", std::string("m"));
        const double rng[2] = {-1.5, 2.5};
        g.putAtt("valid_range", ncDouble, 2, rng);
        int vals[12];
        for (int i = 0; i < 12; ++i) vals[i] = i * i - 5;
        g.putVar(vals);
        const double col[3] = {100., 200., 300.};   // a column written as a hyperslab, from doubles into an int variable
        g.putVar({0, 1}, {3, 1}, col);
        const float sc[4] = {0.5f, 1.5f, 2.5f, 3.5f};
        s.putVar(sc);
        const signed char bb[3] = {-1, 0, 7};
        b.putVar(bb);
        bool threw = false;
        try { f.addVar("wide", ncInt64, dx); } catch (exceptions::NcException const &) { threw = true; }   // needs CDF-5
        CHECK(threw);
    }
    {
        NcFile f(fname, NcFile::read);
        CHECK(f.getDimCount() == 2 && f.getVarCount() == 3);
        std::string t;
        f.getAtt("title").getValues(t);
        CHECK(t == "mini");
        NcVar g = f.getVar("grid");
        long long all[12];
        g.getVar(all);
        for (int i = 0; i < 12; ++i) CHECK(all[i] == ((i % 4 == 1) ? 100 * (i / 4 + 1) : i * i - 5));
        double sub[4];
        g.getVar({1, 2}, {2, 2}, sub);   // rows 1..2, columns 2..3
        CHECK(sub[0] == 31. && sub[1] == 44. && sub[2] == 95. && sub[3] == 116.);
        double rng[2];
        g.getAtt("valid_range").getValues(rng);
        CHECK(rng[0] == -1.5 && rng[1] == 2.5);
        signed char bb[3];
        f.getVar("bytes").getVar(bb);
        CHECK(bb[0] == -1 && bb[1] == 0 && bb[2] == 7);
        bool threw = false;
        try { g.getVar({2, 3}, {2, 1}, sub); } catch (exceptions::NcException const &) { threw = true; }   // beyond the bounds
        CHECK(threw);
        threw = false;
        try { g.putVar(all); } catch (exceptions::NcException const &) { threw = true; }   // read-only
        CHECK(threw);
        CHECK(f.getVar("nope").isNull() && f.getDim("nope").isNull());
    }
    // not a netCDF file / an HDF5 file / a truncated file: exceptions, no crash
    {
        const std::string bad = dir + "/bad.nc";
        FILE *fp = std::fopen(bad.c_str(), "wb");
        std::fputs("\x89HDF\r\n\x1a\n........", fp);
        std::fclose(fp);
        bool threw = false;
        try { NcFile f(bad, NcFile::read); } catch (exceptions::NcException const &e) { threw = std::strstr(e.what(), "HDF5") != nullptr; }
        CHECK(threw);
        fp = std::fopen(bad.c_str(), "wb");
        std::fputs("CDF\x05\0\0", fp);
        std::fclose(fp);
        threw = false;
        try { NcFile f(bad, NcFile::read); } catch (exceptions::NcException const &) { threw = true; }
        CHECK(threw);
        threw = false;
        try { NcFile f(dir + "/does_not_exist.nc", NcFile::read); } catch (exceptions::NcException const &) { threw = true; }
        CHECK(threw);
        std::remove(bad.c_str());
    }
}

static void dump(std::string const &path) {
    NcFile f(path, NcFile::read);
    for (auto const &d : f.getDims()) std::printf("dim %s %zu %d\n", d.first.c_str(), d.second.getSize(), (int)d.second.isUnlimited());
    for (auto const &kv : f.getVars()) {
        NcVar v = kv.second;
        std::printf("var %s %s", v.getName().c_str(), v.getType().getName().c_str());
        size_t n = 1;
        for (int k = 0; k < v.getDimCount(); ++k) { std::printf(" %s", v.getDim(k).getName().c_str()); n *= v.getDim(k).getSize(); }
        std::printf("\n");
        std::vector<double> vals(n);
        v.getVar(vals.data());
        std::printf("values");
        for (double x : vals) std::printf(" %.17g", x);
        std::printf("\n");
    }
}

int main(int argc, char **argv) {
    const std::string cmd = argc > 1 ? argv[1] : "";
    try {
        if (cmd == "selftest" && argc > 2) {
            test_round_trip(argv[2]);
            test_mini_api(argv[2]);
            std::printf("%d failure(s)\n", failures);
            return failures ? 1 : 0;
        }
        if (cmd == "write" && argc > 2) {
            Mat a(fixed_array());
            ibmisc::NcIO ncio(argv[2], NcFile::replace);
            ncio_spsparse(ncio, a, true, "arr1");
            ncio.close();
            return 0;
        }
        if (cmd == "write-classic" && argc > 3) {
            const bool c64 = std::string(argv[3]) == "classic64";
            NcFile f(argv[2], NcFile::replace, c64 ? NcFile::classic64 : NcFile::classic);
            NcDim dn = f.addDim("n", 5), dr = f.addDim("rank", 2);
            NcVar ind = f.addVar("indices", ncInt, {dn, dr}), vals = f.addVar("vals", ncDouble, dn), sh = f.addVar("tag", ncShort, dr);
            const int ix[10] = {1, 2, 3, 3, 4, 5, 1, 2, 0, 0};
            const double vv[5] = {2., 6., 1., -0.25, 0.};
            const short tg[2] = {-7, 300};
            ind.putVar(ix);
            vals.putVar(vv);
            sh.putVar(tg);
            const int shape[2] = {5, 6};
            ind.putAtt("shape", ncInt, 2, shape);
            f.putAtt("history", std::string("written by netcdf_test"));
            return 0;
        }
        if (cmd == "dump" && argc > 2) { dump(argv[2]); return 0; }
    } catch (std::exception const &e) {
        std::printf("exception: %s\n", e.what());
        return 2;
    }
    std::printf("usage: netcdf_test selftest DIR | write PATH | write-classic PATH classic|classic64 | dump PATH\n");
    return 64;
}
