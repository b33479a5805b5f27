#!/bin/bash
# Opcode census of the shipped library (cuobjdump -sass): the tcgen05 / TMEM / bulk-TMA mnemonics per kernel -> profiles/<tag>_sass_census.md
tag=${1:-r02}
so=proud_slam_b200/libproud_b200.so
out=profiles/${tag}_sass_census.md
{
  echo "# ${tag}: SASS census of \`$so\` (\`cuobjdump -sass\`, sm_100a)"
  echo
  echo "| kernel | instructions | UTCHMMA (tcgen05.mma) | UTCQMMA/UTCIMMA | LDTM (tcgen05.ld) | STTM (tcgen05.st) | UBLKCP (bulk TMA) | UTCBAR (tcgen05.commit) | SYNCS (mbarrier) | RED/ATOM |"
  echo "|---|---|---|---|---|---|---|---|---|---|"
  cuobjdump -sass $so | awk '
    /Function : / { if (name != "") print name, n, mma, qmma, ldtm, sttm, blk, bar, syncs, red; name=$3; n=0; mma=0; qmma=0; ldtm=0; sttm=0; blk=0; bar=0; syncs=0; red=0 }
    /^[ \t]+\/\*[0-9a-f]+\*\// { n++ }
    /UTCHMMA/ { mma++ } /UTCQMMA|UTCIMMA/ { qmma++ } /LDTM/ { ldtm++ } /STTM/ { sttm++ } /UBLKCP/ { blk++ } /UTCBAR/ { bar++ } /SYNCS/ { syncs++ } /RED\.|ATOM/ { red++ }
    END { print name, n, mma, qmma, ldtm, sttm, blk, bar, syncs, red }' | while read name n mma qmma ldtm sttm blk bar syncs red; do
      d=$(echo $name | c++filt | sed 's/(.*//; s/^void //; s/pslam:://')
      [ "$n" -gt 0 ] && echo "| \`$d\` | $n | $mma | $qmma | $ldtm | $sttm | $blk | $bar | $syncs | $red |"
    done | sort -t'|' -k4 -n -r
} > $out
head -30 $out
