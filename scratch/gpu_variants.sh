#!/bin/bash
# bench each library variant built by scratch/build_variants.sh
mkdir -p gpurun_out
cp se3et_b200/csrc/libse3et_b200.so /tmp/orig.so
for v in "$@"; do
  cp scratch/variants/$v.so se3et_b200/csrc/libse3et_b200.so
  python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$v.log 2>&1
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/bench_%s.log' % v).read().strip().splitlines()[-1])
    pe = d['roofline']['per_entry_point_ms']
    print(v, 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), {k: pe[k] for k in ('se3et_kpconv_fused', 'se3et_gemm_bf16', 'se3et_gemm_bf16_gnstats', 'se3et_gemm_bf16_gnapply', 'se3et_gemm_grouped_bf16', 'se3et_geo_embed_lookup')})
except Exception as e:
    print(v, 'failed', e)
PY
done
cp /tmp/orig.so se3et_b200/csrc/libse3et_b200.so
