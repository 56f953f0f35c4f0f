#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- tcgen05 3xTF32"; python tools/tc_error_growth.py
echo "--- FFMA"; JB_DISABLE_TC=1 python tools/tc_error_growth.py
