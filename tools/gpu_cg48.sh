#!/bin/bash
for cfg in "X=0" "BRTPE_HALO_CG=2" "BRTPE_HALO_CG=2 BRTPE_HALO_DUAL=0" "X=1"; do
  echo "== $cfg"; env $cfg python tools/halo_prof.py 64 160 160 48 2>&1 | grep -E "^N=|mma.wait|prod.wait|epi.wait"
  env $cfg RES=0 python tools/halo_prof.py 64 160 160 48 2>&1 | grep -E "^N="
done
