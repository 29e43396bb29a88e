#!/bin/bash
# Per-kernel counts of the SASS instructions that prove the Blackwell-native path (tcgen05 MMA kinds, TMA, TMEM loads) in
# the built library:  tools/sass_summary.sh > profiles/r2_sass_summary.txt
LIB=${1:-bayesian_inference_for_nn_b200/libpyesian_b200.so}
echo "# cuobjdump -sass $LIB: per kernel, UTCHMMA = tcgen05.mma kind::f16, UTCIMMA = kind::i8, .2CTA = cta_group::2,"
echo "# UTMALDG = TMA tensor loads, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops"
cuobjdump -sass "$LIB" | awk '
/Function :/ { name=$3; order[++n]=name }
/UTCHMMA\.2CTA/ {h2[name]++; next} /UTCHMMA/ {h1[name]++}
/UTCIMMA\.2CTA/ {i2[name]++; next} /UTCIMMA/ {i1[name]++}
/UTMALDG/ {tma[name]++} /LDTM/ {ldtm[name]++} /UTCBAR/ {cbar[name]++} /SYNCS/ {sy[name]++}
/HMMA\.16816/ {hm[name]++} /LDSM/ {ldsm[name]++} /FFMA2|FADD2/ {f2[name]++}
END { for (k=1;k<=n;k++) { f=order[k]; if (h1[f]+h2[f]+i1[f]+i2[f]+tma[f]+ldtm[f] > 0)
  printf "%-110s UTCHMMA %3d  UTCHMMA.2CTA %3d  UTCIMMA %3d  UTCIMMA.2CTA %3d  UTMALDG %3d  LDTM %3d  UTCBAR %2d  SYNCS %3d  HMMA.16816 %3d  LDSM %2d  FFMA2/FADD2 %3d\n", substr(f,1,110), h1[f], h2[f], i1[f], i2[f], tma[f], ldtm[f], cbar[f], sy[f], hm[f], ldsm[f], f2[f]
  else if (f ~ /k_fs_hmc|k_fs_eval|k_live_sweep_cta|k_scatter/) printf "%-110s (SIMT) FFMA2/FADD2 %3d\n", substr(f,1,110), f2[f] } }' | c++filt | sed 's/CUtensorMap_st, CUtensorMap_st, CUtensorMap_st, CUtensorMap_st, //'
