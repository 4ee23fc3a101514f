#!/bin/bash
# per-source-line instruction / stall-sample shares of an .ncu-rep captured with --import-source on
ncu -i "$1" --page source --csv --print-source cuda,sass > /tmp/_ncu_src.csv 2>/dev/null
python "$(dirname "$0")/prof_lines.py" /tmp/_ncu_src.csv "${2:-30}"
