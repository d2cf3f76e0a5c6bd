#!/bin/bash
# Counts of the SASS mnemonics that prove what each kernel runs on (tcgen05 MMA, TMA, TMEM loads, warp reductions),
# per kernel of the built library -> profiles/r02_sass_summary.txt
set -e
SO=${1:-die_e_b200/libdiee_cuda.so}
OUT=${2:-profiles/r02_sass_summary.txt}
{
echo "cuobjdump -sass $SO (sm_100a); columns: instructions | UTCHMMA (tcgen05.mma) | UTMALDG (TMA tensor load) | LDTM (tcgen05.ld) | UTCBAR/UTCATOMSWS (tcgen05 commit/alloc) | SYNCS (mbarrier) | REDUX (redux.sync) | POPC | VOTE | MATCH | ATOM/RED | kernel"
cuobjdump -sass "$SO" | awk '
/Function :/ { if (name != "") printf "%7d %5d %5d %5d %5d %5d %5d %5d %5d %5d %5d  %s\n", n, mma, tma, ldtm, utc, syncs, redux, popc, vote, match_, atom, name;
               name=$3; n=0; mma=0; tma=0; ldtm=0; utc=0; syncs=0; redux=0; popc=0; vote=0; match_=0; atom=0 }
/^ +\/\*[0-9a-f]+\*\// { n++;
  if ($0 ~ /UTCHMMA|UTCQMMA|UTCIMMA/) mma++;
  if ($0 ~ /UTMALDG/) tma++;
  if ($0 ~ /LDTM/) ldtm++;
  if ($0 ~ /UTCBAR|UTCATOMSWS/) utc++;
  if ($0 ~ /SYNCS/) syncs++;
  if ($0 ~ /REDUX/) redux++;
  if ($0 ~ /POPC/) popc++;
  if ($0 ~ /VOTE/) vote++;
  if ($0 ~ /MATCH/) match_++;
  if ($0 ~ / ATOM| RED\./) atom++; }
END { if (name != "") printf "%7d %5d %5d %5d %5d %5d %5d %5d %5d %5d %5d  %s\n", n, mma, tma, ldtm, utc, syncs, redux, popc, vote, match_, atom, name }' | c++filt | sort -k12
} > "$OUT"
wc -l "$OUT"
