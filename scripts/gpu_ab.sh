#!/bin/bash
# A/B of prebuilt library variants: libreacher_b200_v*.so
cp reacherdistilation_b200/libreacher_b200.so /tmp/orig.so
for f in reacherdistilation_b200/libreacher_b200_v*.so; do
  cp $f reacherdistilation_b200/libreacher_b200.so
  echo "== $f"
  python bench.py --steps 50 --warmup 5 --quick 2>&1 | grep -o '"value": [0-9.]*' | head -1
  python -m pytest tests/test_policy_gpu.py -m gpu -q -x -s 2>&1 | grep -E "fused rollout|passed|failed|n=4097 nout=2"
done
cp /tmp/orig.so reacherdistilation_b200/libreacher_b200.so
