for f in default mb9 mb11 mb12; do
  if [ $f = default ]; then unset WARP_B200_LIB; else export WARP_B200_LIB=$PWD/warp_b200/lib/variants/$f.so; fi
  python bench.py --steps 6 --warmup 3 --no-extra --no-cpu-baseline --skip parity 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$f value %.1f M kernel %.2f ms' % (d['value']/1e6, d['roofline']['launch_ms']))"
done
