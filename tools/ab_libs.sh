# same-box comparison of library builds / switches on single layers: ab_libs.sh "shape..." -- "ENV=.. ENV=.." ...
SH="$1"; shift
for cfg in "$@"; do
  echo "== $cfg"; env $cfg python tools/halo_prof.py $SH 2>&1 | grep -E "^N="
done
