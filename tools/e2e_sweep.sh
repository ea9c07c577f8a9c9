#!/bin/bash
for s in 512 1024 1376 2048 4096; do
  echo -n "host stage $s: "; ARX_HOST_STAGE=$s timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python tools/benchsum.py | cut -c1-60
done
echo -n "default: "; timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | python tools/benchsum.py | cut -c1-60
