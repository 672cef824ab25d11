// NetCDF I/O of a COO array -- same names, arguments and file layout as the reference's slib/spsparse/netcdf.hpp:
//   dimensions  <v>.size (number of entries), <v>.rank
//   variables   <v>.info    int64 scalar, attribute "shape" uint64[rank]          (netcdf.hpp:99-100)
//               <v>.indices int64 [<v>.size, <v>.rank]                           (netcdf.hpp:102)
//               <v>.vals    double [<v>.size]                                    (netcdf.hpp:103)
// ncio_spsparse (netcdf.hpp:86-138) defines or checks these and registers the transfer with the NcIO, which runs it at
// close(); nc_write_spsparse / nc_read_spsparse (netcdf.hpp:23-43, 52-76) are the transfers.  The reference moves one
// entry per putVar/getVar call; here entries move in blocks of NC_BLOCK through one contiguous buffer per variable --
// same file contents, same order of add() calls on the reader's side.
//
// Written against the subset of the netCDF C++4 / ibmisc interfaces that the reference uses, so it compiles with the real
// libraries (their include directories first) or, where there is no netCDF -- this image -- with the minimal stand-ins under
// include/spsparse_b200/mini_netcdf/ (-I that directory: <netcdf>, <ibmisc/netcdf.hpp>; classic-format files, see there).
#pragma once
#include <algorithm>
#include <array>
#include <functional>
#include <string>
#include <vector>

#include <ibmisc/netcdf.hpp>

#include "base.hpp"

namespace spsparse {

namespace b200 {
constexpr size_t NC_BLOCK = size_t(1) << 20;  // entries per putVar / getVar call
}

template <class ArrayT>
void nc_write_spsparse(netCDF::NcGroup *nc, ArrayT *A, std::string const &vname) {
    constexpr size_t R = (size_t)ArrayT::rank;
    netCDF::NcVar indices_v = nc->getVar(vname + ".indices");
    netCDF::NcVar vals_v = nc->getVar(vname + ".vals");
    std::vector<long long> ibuf;
    std::vector<double> vbuf;
    ibuf.reserve(std::min(A->size(), b200::NC_BLOCK) * R);
    vbuf.reserve(std::min(A->size(), b200::NC_BLOCK));
    size_t start = 0;
    auto flush = [&]() {
        if (vbuf.empty()) return;
        indices_v.putVar(std::vector<size_t>{start, 0}, std::vector<size_t>{vbuf.size(), R}, ibuf.data());
        vals_v.putVar(std::vector<size_t>{start}, std::vector<size_t>{vbuf.size()}, vbuf.data());
        start += vbuf.size();
        ibuf.clear();
        vbuf.clear();
    };
    for (auto ii = A->begin(); ii != A->end(); ++ii) {
        for (size_t k = 0; k < R; ++k) ibuf.push_back((long long)ii.index((int)k));
        vbuf.push_back((double)ii.val());
        if (vbuf.size() == b200::NC_BLOCK) flush();
    }
    flush();
}

template <class AccumulatorT>
void nc_read_spsparse(netCDF::NcGroup *nc, AccumulatorT *A, std::string const &vname) {
    constexpr size_t R = (size_t)AccumulatorT::rank;
    netCDF::NcVar indices_v = nc->getVar(vname + ".indices");
    netCDF::NcVar vals_v = nc->getVar(vname + ".vals");
    const size_t size = vals_v.getDim(0).getSize();  // number of stored entries
    std::vector<long long> ibuf(std::min(size, b200::NC_BLOCK) * R);
    std::vector<double> vbuf(std::min(size, b200::NC_BLOCK));
    std::array<typename AccumulatorT::index_type, AccumulatorT::rank> index;
    for (size_t start = 0; start < size; start += b200::NC_BLOCK) {
        const size_t cnt = std::min(b200::NC_BLOCK, size - start);
        indices_v.getVar(std::vector<size_t>{start, 0}, std::vector<size_t>{cnt, R}, ibuf.data());
        vals_v.getVar(std::vector<size_t>{start}, std::vector<size_t>{cnt}, vbuf.data());
        for (size_t t = 0; t < cnt; ++t) {
            for (size_t k = 0; k < R; ++k) index[k] = (typename AccumulatorT::index_type)ibuf[t * R + k];
            A->add(index, (typename AccumulatorT::val_type)vbuf[t]);
        }
    }
}

template <class ArrayT>
void ncio_spsparse(ibmisc::NcIO &ncio, ArrayT &A, bool alloc, std::string const &vname) {
    std::vector<std::string> const dim_names({vname + ".size", vname + ".rank"});
    if (ncio.rw == 'w') {
        std::vector<netCDF::NcDim> dims = ibmisc::get_or_add_dims(ncio, dim_names, {A.size(), (size_t)A.rank});
        netCDF::NcVar info_v = ibmisc::get_or_add_var(ncio, vname + ".info", netCDF::ncInt64, {});
        std::array<unsigned long long, ArrayT::rank> shape;
        for (int k = 0; k < ArrayT::rank; ++k) shape[k] = (unsigned long long)A.shape[k];
        info_v.putAtt("shape", netCDF::ncUint64, (size_t)ArrayT::rank, &shape[0]);
        ibmisc::get_or_add_var(ncio, vname + ".indices", netCDF::ncInt64, dims);
        ibmisc::get_or_add_var(ncio, vname + ".vals", netCDF::ncDouble, {dims[0]});
        ncio += std::bind(&nc_write_spsparse<ArrayT>, ncio.nc, &A, vname);
    } else {
        ibmisc::get_dims(ncio, dim_names);  // both must exist
        netCDF::NcVar info_v = ncio.nc->getVar(vname + ".info");
        auto shape_a = info_v.getAtt("shape");
        // the rank in the file must be the array's
        const size_t rank = shape_a.getAttLength();
        if (rank != (size_t)ArrayT::rank)
            (*spsparse_error)(-1, "Trying to read NetCDF sparse array of rank %ld into SpSparse array of rank %d", (long)rank, (int)ArrayT::rank);
        if (alloc) {  // take the shape from the file and make room for the stored entries
            std::array<unsigned long long, ArrayT::rank> file_shape;
            shape_a.getValues(&file_shape[0]);
            std::array<size_t, ArrayT::rank> shape;
            for (int k = 0; k < ArrayT::rank; ++k) shape[k] = (size_t)file_shape[k];
            A.clear();
            A.set_shape(shape);
            A.reserve(ncio.nc->getDim(vname + ".size").getSize());
        }
        ncio += std::bind(&nc_read_spsparse<ArrayT>, ncio.nc, &A, vname);
    }
}

}  // namespace spsparse
