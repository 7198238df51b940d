#!/bin/bash
# parity tests under a tuning-env setting:  tools/variants_test.sh "VAR=val ..."
for v in "$@"; do
  echo "== pytest -m gpu with: $v"
  env $v timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
done
