#!/bin/bash
# A/B of library variants: tools/ab.sh tag1 tag2 ...   (runs bench at quarter scale for each libbdf_<tag>.so; tag "ws" = the
# shipped library with the persistent warp-specialised row kernel switched on, BDF_ROWS_WS=1)
for t in "$@"; do
  lib=$PWD/bayesiandatafusion.jl_b200/libbdf_$t.so; ws=0
  if [ "$t" = ws ]; then lib=$PWD/bayesiandatafusion.jl_b200/libbdf_b200.so; ws=1; fi
  BDF_ROWS_WS=$ws BDF_B200_LIB=$lib timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu --scale ${SCALE:-0.25} $BENCH_ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', round(d['ms_per_step'],2), 'ms/sweep  frac', round(d['roofline']['frac'],3), {k:round(v,2) for k,v in d['roofline']['ms_per_launch'].items()})"
done
