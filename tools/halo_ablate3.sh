# ablation of conv_halo_kernel parts (BRTPE_HALO_DBG bits: 1 no MMA, 2 no epilogue, 4 no weight loads, 8 no activation loads)
SHAPE=${SHAPE:-"64 160 160 48"}
for v in "" "BRTPE_HALO_DBG=2" "BRTPE_HALO_DBG=8" "BRTPE_HALO_DBG=10" "BRTPE_HALO_TMA_OUT=1" "BRTPE_HALO_TMA_OUT=0" "RES=0" "RES=0 BRTPE_HALO_DBG=8"; do
  echo "== $v"; env $v python tools/halo_prof.py $SHAPE 2>&1 | grep -E "^N=|mma.wait|epi.wait"
done
