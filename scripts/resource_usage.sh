#!/bin/bash
# Static evidence from the built library (no GPU needed): per-kernel registers / stack / static shared memory and the
# SASS mnemonics that show which hardware paths the hot kernels use (tcgen05 MMA = UTC*MMA, TMA = UTMALDG / UBLKCP,
# mbarrier = SYNCS, FP64 FMA = DFMA, FP64 tensor op = DMMA, tensor-memory loads = LDTM).  Usage: scripts/resource_usage.sh > profiles/rNN_static_resources.md
set -e
SO=${1:-gemma_b200/csrc/libgemma_b200.so}
echo "# Static resource usage and SASS evidence: $(basename $SO), built $(date -u -r $SO +%Y-%m-%dT%H:%MZ), $(nvcc --version | grep release | sed 's/.*release //')"
echo
echo "| kernel | regs | stack B | static smem B |"
echo "|---|---|---|---|"
cuobjdump -res-usage $SO 2>/dev/null | tr '\n' ' ' | sed 's/Function /\n/g' | grep REG | while read -r line; do
  name=$(echo "$line" | cut -d: -f1 | c++filt | sed 's/(.*//; s/^void //')
  reg=$(echo "$line" | grep -o "REG:[0-9]*" | cut -d: -f2); st=$(echo "$line" | grep -o "STACK:[0-9]*" | cut -d: -f2)
  sh=$(echo "$line" | grep -o "SHARED:[0-9]*" | cut -d: -f2)
  echo "| \`$name\` | $reg | $st | $sh |"
done
echo
echo "## SASS mnemonic counts of the hot kernels"
echo
echo "| kernel | UTC*MMA (tcgen05.mma) | UTMALDG (TMA tensor load) | UBLKCP (TMA bulk copy) | LDTM (tcgen05.ld) | SYNCS (mbarrier) | DFMA | DMMA (FP64 mma.sync) | IMMA/HMMA (legacy mma.sync) |"
echo "|---|---|---|---|---|---|---|---|---|"
for k in i8_gemm_kernel i8_gemm_pair_kernel lmm_assoc_v2_kernelILi1E lmm_common_kernelILi1E lmm_assoc_kernelILi1E mv_assoc_kernelILi1E dgemm_kernelILb0E kin_miss_fix_kernel; do
  sass=$(cuobjdump -sass -fun $(cuobjdump -res-usage $SO 2>/dev/null | grep -o "_ZN2gb[A-Za-z0-9_]*" | grep "$k" | head -1) $SO 2>/dev/null)
  c() { echo "$sass" | grep -c "$1" || true; }
  echo "| \`$k\` | $(c 'UTC[A-Z]*MMA') | $(c UTMALDG) | $(c UBLKCP) | $(c LDTM) | $(c SYNCS) | $(c DFMA) | $(c DMMA) | $(c ' [IH]MMA\.') |"
done
