#!/bin/bash
# tools/mk_variant.sh <name> [extra nvcc flags for agb_walk.cu]: dev_libs/libagb200_<name>.so = the in-tree objects with a
# differently compiled walk (tuning experiments; select with AGB200_LIB=<path>).  Run astrogenesis2.0_b200/build.py first.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=astrogenesis2.0_b200/csrc
mkdir -p dev_libs
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $C/agb_walk.cu -o dev_libs/agb_walk_$name.o
nvcc -shared -o dev_libs/libagb200_$name.so $C/agb_api.o $C/agb_build.o $C/agb_density.o $C/agb_integrate.o $C/agb_multi.o $C/agb_extended.o dev_libs/agb_walk_$name.o -lcudart -lpthread -ldl
rm -f dev_libs/agb_walk_$name.o
ls -la dev_libs/libagb200_$name.so
