#!/bin/bash
# chunk-size sweep of bench.py (run on the GPU box)
for c in 0 2048 1024 512 256; do
  echo -n "chunk $c: "; timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --chunk $c | python tools/benchsum.py
done
