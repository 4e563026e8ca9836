#!/bin/bash
for G in 0 1; do
  echo "== MS_PANEL_GEMM=$G"
  MS_PANEL_GEMM=$G MS_PROFILE_UPDATE=1 python tools/gpu_check.py --stage mid 2>&1 | python tools/fmt_check.py
  MS_PANEL_GEMM=$G MS_PROFILE_UPDATE=1 python tools/gpu_check.py --stage big 2>&1 | python tools/fmt_check.py
done
