#!/bin/bash
# Round-2 A/B of the lockstep per-SNP kernel's occupancy point (profiles/r01_static_resources.md): builds the variant
# libraries next to the default one (only lmm_kernel.cu differs) and prints the gpurun command that times all of them
# back to back on one box through GB200_LIB.  Build here (no GPU needed); run the printed command.
set -e
cd "$(dirname "$0")/../gemma_b200/csrc"
make -s -j8
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr"
mkdir -p variants
specs=()
for v in "w6:-DGB_V2_WARPS=6" "cta1:-DGB_V2_CTAS2_MAXNC=0" "w4:-DGB_V2_WARPS=4"; do
  tag=${v%%:*}; def=${v#*:}
  /usr/local/cuda/bin/nvcc $FLAGS $def -c lmm_kernel.cu -o variants/lmm_kernel_$tag.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libgemma_b200_$tag.so \
      api.o variants/lmm_kernel_$tag.o mvlmm_kernel.o dgemm.o geno.o eigh.o i8gemm_sm100.o -lcusolver -lcublas
  specs+=("\"$tag|--steps 5 --warmup 3|GB200_LIB=gemma_b200/csrc/variants/libgemma_b200_$tag.so\"")
done
echo "gpurun --timeout 1500 -- 'scripts/bench_ab.sh \"base|--steps 5 --warmup 3\" ${specs[*]}'"
