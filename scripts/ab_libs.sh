# Same-box A/B of two builds of libnerf_b200.so (box-to-box spread of the kernel timings is +-3 %, larger than most single changes):
# usage: bash scripts/ab_libs.sh <alt.so> [rounds]   -> alternates the shipped build (A) and <alt.so> (B) under scripts/abl_probe.py
ALT=$1; ROUNDS=${2:-2}
LIB=nerf_pytorch_paeng_b200/libnerf_b200.so
cp $LIB /tmp/libA.so; cp $ALT /tmp/libB.so
for r in $(seq $ROUNDS); do
  for v in A B; do
    cp /tmp/lib$v.so $LIB
    echo "== build $v round $r"
    timeout 100 python scripts/abl_probe.py 2>/dev/null | head -2 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: round(v, 4) for k, v in d.items() if k.endswith('_ms') or k == 'points'})"
  done
done
cp /tmp/libA.so $LIB
