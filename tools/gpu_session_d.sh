#!/bin/bash
KEMR_MMA_PAIR=0 timeout 300 python tools/diag_pair_count.py 2>&1 | tail -12
KEMR_MMA_PAIR=1 timeout 300 python tools/diag_pair_count.py 2>&1 | tail -12
