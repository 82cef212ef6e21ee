#!/bin/bash
# Tuning sweep on the GPU box: bench the tracking kernel built with different compile-time knobs.
# Variant libraries are built HERE (no GPU needed) by scripts/build_variants.sh and travel as .so files.
OUT=gpurun_out/${1:-sweep}
mkdir -p $OUT
for sfx in "" $(ls xtrack_b200/libxtb200_*.so 2>/dev/null | sed -E 's/.*libxtb200(_[^.]*)\.so/\1/'); do
  for mode in "" "--fma"; do
    XTB_LIB_SUFFIX=$sfx timeout 300 python bench.py --steps 3 --warmup 1 --turns 10 --no-cpu-baseline --quick $mode \
        > $OUT/bench${sfx}${mode}.json 2>> $OUT/err.log
    python - <<PY
import json
try:
    d = json.load(open("$OUT/bench${sfx}${mode}.json"))
    print("variant[$sfx] mode[$mode] PET/s %.3e frac %.3f kernel_ms %.1f" % (d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_step"]))
except Exception as e:
    print("variant[$sfx] mode[$mode] FAILED", e)
PY
  done
done
