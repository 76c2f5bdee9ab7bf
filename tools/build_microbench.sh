#!/bin/bash
# builds the sm_100a microbenchmarks next to their sources (run them on the B200 with gpurun)
cd "$(dirname "$0")"
for f in ffma2_bench ffma2_bench2 gather_bench; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o $f.bin $f.cu || exit 1
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pair_bench.bin pair_bench.cu || exit 1
# the library's radix sort against std::stable_sort and CUB (links the library for its allocator)
P=../parallel-reverb-raytracer_b200
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o sort_test.bin sort_test.cu -L$P -lrayverb_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../parallel-reverb-raytracer_b200' || exit 1
