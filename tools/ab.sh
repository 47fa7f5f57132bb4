#!/bin/bash
# A/B of library variants (multivolumes_b200/libmv_var_<name>.so, built with `make BUILD=_build_<name>
# OUT=../libmv_var_<name>.so EXTRA=-D...`) on the per-pass device times: tools/ab.sh "base spec" "cfg2 cfg4"
VARS=${1:-base}; WLS=${2:-cfg2}
for wl in $WLS; do for v in $VARS; do
  lib=multivolumes_b200/libmv_var_$v.so; [ "$v" = base ] && lib=multivolumes_b200/libmv_b200.so
  MV_B200_LIB=$PWD/$lib MV_NOSTATS=1 python tools/pass_times.py $wl ${FRAMES:-60} | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('$wl %-8s L %.4f V %.4f OIT %.4f post %.4f total %.4f' % ('$v', d['ray_march_light'], d['ray_march_view'], d['resolve_oit'], d['postprocess'], d['total']))"
done; done
