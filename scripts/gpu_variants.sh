#!/bin/bash
# kernel times of every lib variant under supernormal_b200/lib/variants (built by scripts/build_variant.sh)
cp supernormal_b200/lib/libsnb200.so /tmp/libsnb200_shipped.so
for v in "$@"; do
  echo "== variant $v"
  cp supernormal_b200/lib/variants/libsnb200_$v.so supernormal_b200/lib/libsnb200.so
  python scripts/kernel_times.py 15 1000 4800 2>&1 | grep iter | cut -c1-420
done
cp /tmp/libsnb200_shipped.so supernormal_b200/lib/libsnb200.so
