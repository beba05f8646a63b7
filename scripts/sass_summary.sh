#!/bin/bash
# SASS evidence of the Blackwell-native paths: counts of the tcgen05 / TMEM / TMA / packed-FP32 mnemonics per object of
# libpsgla_b200.so (build the library first).   bash scripts/sass_summary.sh > profiles/r02_sass_summary.txt
cd "$(dirname "$0")/../psgla-for-posterior-sampling_b200/build" || exit 1
echo "# cuobjdump -sass <object> | grep -c <mnemonic>, per object of libpsgla_b200.so ($(nvcc --version | tail -2 | head -1), -gencode arch=compute_100a,code=sm_100a)"
echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy,"
echo "# UTCBAR = tcgen05.commit, FFMA2 = fma.rn.f32x2, HMMA = legacy mma.sync (must be 0)"
printf "%-22s %8s %13s %6s %6s %8s %8s %7s %7s %7s %6s\n" object UTCHMMA UTCHMMA.2CTA LDTM STTM UTMALDG UTMASTG UTCBAR UBLKCP FFMA2 HMMA
for o in *.o; do
  t=$(mktemp); cuobjdump -sass "$o" > "$t" 2>/dev/null
  printf "%-22s %8d %13d %6d %6d %8d %8d %7d %7d %7d %6d\n" "$o" "$(grep -c 'UTCHMMA' "$t")" "$(grep -c 'UTCHMMA.2CTA' "$t")" "$(grep -c 'LDTM' "$t")" \
    "$(grep -c 'STTM' "$t")" "$(grep -c 'UTMALDG' "$t")" "$(grep -c 'UTMASTG' "$t")" "$(grep -c 'UTCBAR' "$t")" "$(grep -c 'UBLKCP' "$t")" \
    "$(grep -c 'FFMA2' "$t")" "$(grep -cE '(^|[^A-Z])HMMA' "$t")"
  rm -f "$t"
done
