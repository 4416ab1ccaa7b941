#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG=.. -DFLAG2=.." : builds morb_slam_b200/lib/variants/liborb_b200_NAME.so with extra compile flags for
# orb_extract.cu (the extraction kernels); the other objects are the default build's. Select it with ORB_B200_LIB=<path> (A/B measurements).
set -e
cd "$(dirname "$0")/.."
C=morb_slam_b200/csrc
mkdir -p morb_slam_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 -Iinclude -I$C $2 -c -o /tmp/orb_extract_$1.o $C/orb_extract.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o morb_slam_b200/lib/variants/liborb_b200_$1.so /tmp/orb_extract_$1.o $C/orb_stereo.o $C/orb_knn.o $C/orb_match.o $C/orb_bow.o $C/orb_fisheye.o $C/orb_serialize.o $C/orb_mapping.o
echo built morb_slam_b200/lib/variants/liborb_b200_$1.so
