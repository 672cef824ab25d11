#!/bin/bash
# round 2, call 46 (1 GPU): the NetCDF round trip of the Python mirror
timeout 60 python -m pytest tests/test_gpu_dense_ops.py -x -q -p no:cacheprovider -k netcdf 2>&1 | tail -n 5
