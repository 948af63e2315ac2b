# A/B of a halo-kernel switch on the same box: usage ab_halo.sh VAR "shapes..."
VAR=$1; shift
for v in 1 0; do
  echo "== $VAR=$v"; env $VAR=$v python tools/halo_prof.py "$@" 2>&1 | grep -E "^N="
done
