#!/bin/bash
# Per-kernel SASS evidence of the tcgen05 / TMEM / TMA instructions ("commit one").  Runs without a GPU.
# usage: scripts/sass_evidence.sh > profiles/rNN_sass_counts.txt
SO=${1:-dlwp_benchmark_b200/libspectral_b200.so}
echo "# cuobjdump -sass $SO  ($(date -u +%FT%TZ), $(git rev-parse --short HEAD 2>/dev/null))"
echo "# columns: UTCHMMA (tcgen05.mma kind::tf32) | UTCBAR (tcgen05.commit) | LDTM (tcgen05.ld) | UTMALDG (TMA load) | UTMASTG (TMA store) | SYNCS (mbarrier) | FFMA2 | FFMA | kernel"
cuobjdump -sass "$SO" | awk '
  /Function :/ { if (name != "") printf "%7d %6d %5d %7d %7d %6d %6d %6d  %s\n", mma, bar, ldtm, tmal, tmas, syncs, ffma2, ffma, name;
                 name=$3; mma=bar=ldtm=tmal=tmas=syncs=ffma2=ffma=0; next }
  /UTCHMMA/ {mma++} /UTCBAR/ {bar++} /LDTM/ {ldtm++} /UTMALDG/ {tmal++} /UTMASTG/ {tmas++} /SYNCS/ {syncs++}
  /FFMA2/ {ffma2++} / FFMA / {ffma++}
  END { printf "%7d %6d %5d %7d %7d %6d %6d %6d  %s\n", mma, bar, ldtm, tmal, tmas, syncs, ffma2, ffma, name }' | while read -r a b c d e f g h n; do
    printf "%7s %6s %5s %7s %7s %6s %6s %6s  %s\n" "$a" "$b" "$c" "$d" "$e" "$f" "$g" "$h" "$(echo "$n" | c++filt | cut -c1-110)"
  done | sort -k9
