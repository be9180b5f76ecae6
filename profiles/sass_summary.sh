#!/bin/bash
# SASS evidence for DESIGN.md §4: tensor-core / TMA / TMEM instruction counts per kernel of the built library,
# plus the number of ELECT/BRA.U.ANY issue wrappers (0 = tcgen05.mma and TMA issue straight from uniform registers).
so=${1:-neutts_air_b200/libneutts_b200.so}
cuobjdump -sass "$so" | awk '
  /Function : /{f=$3}
  /UTCHMMA/{mma[f]++} /UTCQMMA|UTCOMMA/{mma8[f]++} /LDTM/{ldtm[f]++} /UTMALDG/{tma[f]++} /UBLKCP/{blk[f]++}
  /HMMA\.16816/{hmma[f]++} /LDSM/{ldsm[f]++} /UTCBAR/{bar[f]++} /SYNCS/{syncs[f]++} /BRA\.U\.ANY/{wrap[f]++} /ELECT/{el[f]++}
  {n[f]++}
  END{printf "%-86s %7s %6s %5s %6s %6s %6s %5s %6s %6s %6s\n","kernel","instr","UTCHMMA","LDTM","UTMALDG","UBLKCP","HMMA","LDSM","UTCBAR","SYNCS","wrappers";
      for(k in n) if (mma[k]+tma[k]+blk[k]+hmma[k]>0) printf "%-86s %7d %6d %5d %6d %6d %6d %5d %6d %6d %6d\n",substr(k,1,86),n[k]/2,mma[k],ldtm[k],tma[k],blk[k],hmma[k],ldsm[k],bar[k],syncs[k],wrap[k]}' | sort
