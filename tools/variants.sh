#!/bin/bash
# Time one configuration with the shipped library and with every tuning variant under assimulo_b200/_variants/
# (built by `python assimulo_b200/build.py --variant NAME -D FLAG=...`):   bash tools/variants.sh radau gyro ...
for cfg in "$@"; do
  echo "== $cfg: shipped"; python tools/run_cfg.py $cfg 3 2>&1 | tail -2
  for v in assimulo_b200/_variants/*/; do
    echo "== $cfg: $(basename $v)"; LD_LIBRARY_PATH=$v python tools/run_cfg.py $cfg 3 2>&1 | tail -2
  done
done
