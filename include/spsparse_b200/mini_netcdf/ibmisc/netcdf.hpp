// <ibmisc/netcdf.hpp> -- the part of ibmisc's NetCDF helper that spsparse's I/O uses (reference slib/spsparse/netcdf.hpp:86-138,
// tests/test_netcdf.cpp:62-85): NcIO, a file handle that remembers whether it reads or writes and a list of deferred
// actions, get_or_add_dims / get_dims / get_or_add_var.  ibmisc itself is an un-vendored dependency of the reference
// (SURVEY.md section 2); this is written from its call sites, for machines that have neither ibmisc nor netCDF
// (together with the minimal <netcdf> next to it).
//
// Protocol (as the reference's call sites use it): in 'w' mode a caller first DEFINES dimensions and variables, and
// registers with `ncio += fn` the function that later writes the data; in 'r' mode it reads sizes and attributes
// immediately and registers the function that reads the data.  close() (or the destructor) runs the registered
// functions in order and then closes the file.
#ifndef SPSPARSE_B200_MINI_IBMISC_NETCDF_HPP
#define SPSPARSE_B200_MINI_IBMISC_NETCDF_HPP

#include <functional>
#include <netcdf>
#include <string>
#include <vector>

namespace ibmisc {

class NcIO {
    netCDF::NcFile file_;
    std::vector<std::function<void()>> todo_;
    bool open_;

public:
    netCDF::NcGroup *const nc;  // the open file (the reference hands this to nc_read_spsparse / nc_write_spsparse)
    const char rw;              // 'r' or 'w'
    const bool define;          // 'w': variables may be defined

    NcIO(const std::string &path, netCDF::NcFile::FileMode mode)
        : file_(path, mode), open_(true), nc(&file_), rw(mode == netCDF::NcFile::read ? 'r' : 'w'), define(rw == 'w') {}
    NcIO(const std::string &path, netCDF::NcFile::FileMode mode, netCDF::NcFile::FileFormat fmt)
        : file_(path, mode, fmt), open_(true), nc(&file_), rw(mode == netCDF::NcFile::read ? 'r' : 'w'), define(rw == 'w') {}
    NcIO(const NcIO &) = delete;
    NcIO &operator=(const NcIO &) = delete;
    ~NcIO() {
        try { close(); } catch (...) {}
    }
    void operator+=(std::function<void()> const &fn) { todo_.push_back(fn); }
    // runs the deferred reads / writes without closing
    void flush() {
        std::vector<std::function<void()>> run;
        run.swap(todo_);
        for (auto &fn : run) fn();
    }
    void close() {
        if (!open_) return;
        open_ = false;
        flush();
        file_.close();
    }
};

// 'w': the named dimensions, created with the given sizes when missing (sizes of existing ones must agree);
// 'r': the named dimensions as they are in the file
inline std::vector<netCDF::NcDim> get_or_add_dims(NcIO &ncio, std::vector<std::string> const &names, std::vector<size_t> const &sizes) {
    if (names.size() != sizes.size()) throw netCDF::exceptions::NcException("get_or_add_dims: names and sizes differ in length");
    std::vector<netCDF::NcDim> out;
    for (size_t i = 0; i < names.size(); ++i) {
        netCDF::NcDim d = ncio.nc->getDim(names[i]);
        if (d.isNull()) {
            if (ncio.rw != 'w') throw netCDF::exceptions::NcException("dimension " + names[i] + " not found");
            d = ncio.nc->addDim(names[i], sizes[i]);
        } else if (ncio.rw == 'w' && d.getSize() != sizes[i]) {
            throw netCDF::exceptions::NcException("dimension " + names[i] + " exists with a different size");
        }
        out.push_back(d);
    }
    return out;
}
inline std::vector<netCDF::NcDim> get_dims(NcIO &ncio, std::vector<std::string> const &names) {
    std::vector<netCDF::NcDim> out;
    for (auto const &nm : names) {
        netCDF::NcDim d = ncio.nc->getDim(nm);
        if (d.isNull()) throw netCDF::exceptions::NcException("dimension " + nm + " not found");
        out.push_back(d);
    }
    return out;
}
// the named variable; in 'w' mode it is defined when missing
inline netCDF::NcVar get_or_add_var(NcIO &ncio, std::string const &name, netCDF::NcType const &type, std::vector<netCDF::NcDim> const &dims) {
    netCDF::NcVar v = ncio.nc->getVar(name);
    if (v.isNull()) {
        if (ncio.rw != 'w') throw netCDF::exceptions::NcException("variable " + name + " not found");
        v = ncio.nc->addVar(name, type, dims);
    }
    return v;
}

}  // namespace ibmisc
#endif
