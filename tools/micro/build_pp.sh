#!/bin/bash
# builds tools/micro/pp_bench (timing) and tools/micro/pp_bench_trace (clock-stamp traces); extra nvcc flags: "$@"
set -e
cd "$(dirname "$0")/../.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DSLMGS_PACKED_F32X2 -DSLMGS_TW_PRODUCTS -lineinfo -I slmsuite_b200/csrc -I tools/micro"
nvcc $FLAGS "$@" -o tools/micro/pp_bench tools/micro/pp_bench.cu &
nvcc $FLAGS "$@" -DSLMGS_PP_TRACE -DSLMGS_PP_TRACE_TID=37 -o tools/micro/pp_bench_trace tools/micro/pp_bench.cu &
wait
