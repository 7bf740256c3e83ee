#!/bin/bash
# Census of the Blackwell-native SASS mnemonics per kernel object (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
# LDTM/STTM, TMA / bulk copies -> UTMALDG / UBLKCP, cp.async -> LDGSTS, legacy mma.sync -> HMMA).
#   tools/sass_census.sh > profiles/r02_sass_census.txt
cd "$(dirname "$0")/../gator_b200/_C/obj" || exit 1
printf '%-26s %8s %6s %6s %7s %8s %7s %6s %9s\n' object UTCHMMA LDTM STTM UBLKCP UTMALDG LDGSTS HMMA SYNCS
for o in *.o; do
  s=$(cuobjdump -sass "$o")
  c() { echo "$s" | grep -c "$1"; }
  printf '%-26s %8d %6d %6d %7d %8d %7d %6d %9d\n' "$o" "$(c 'UTCHMMA')" "$(c 'LDTM')" "$(c 'STTM')" "$(c 'UBLKCP')" "$(c 'UTMALDG')" "$(c 'LDGSTS')" "$(c ' HMMA')" "$(c 'SYNCS')"
done
echo
echo "per kernel (tensor-core kernels):"
for o in mdr_attn2_umma.o mdr_chain2_umma.o gat_chain2_umma.o umma_gemm_wide.o smpl_skin_umma.o umma_gemm.o; do
  cuobjdump -sass "$o" | awk -v obj="$o" '/Function :/{name=$NF} /UTCHMMA/{m[name]++} /LDTM/{l[name]++} /STTM/{s[name]++} /UBLKCP/{b[name]++} /UTCHMMA.*tmem\[UR[0-9]+\], gdesc/{ts[name]++} END{for(n in m) printf "  %-20s UTCHMMA %4d (A operand in tensor memory: %4d)  LDTM %3d  STTM %3d  UBLKCP %2d  %s\n", obj, m[n], ts[n], l[n], s[n], b[n], substr(n,1,90)}'
done
