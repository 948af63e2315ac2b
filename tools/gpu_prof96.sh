#!/bin/bash
python tools/halo_prof.py 64 80 80 96 64 40 40 192 64 160 160 48 2>&1 | tail -60
echo "== no residual"
RES=0 python tools/halo_prof.py 64 80 80 96 64 40 40 192 2>&1 | tail -40
for cfg in "BRTPE_HALO_A_STAGES=2" "BRTPE_HALO_A_STAGES=3" "BRTPE_HALO_A_STAGES=4" "BRTPE_HALO_TMA_OUT=0" "BRTPE_HALO_RES_PREFETCH=1" "BRTPE_HALO_CG=1"; do
  echo "== $cfg"; env $cfg python tools/halo_prof.py 64 80 80 96 2>&1 | grep -E "^N=|ms"
done
