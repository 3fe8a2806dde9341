#!/bin/bash
set -u
O=gpurun_out
python tools/trace_fused.py --screen --no-onehot --shape=1024,64,1024,512 > $O/r2_trace_k512_d64.txt 2>&1
python tools/trace_fused.py --screen --no-onehot --shape=1024,64,1024,1024 > $O/r2_trace_k1024_d64.txt 2>&1
python tools/trace_fused.py --screen --no-onehot --shape=1024,128,1024,512 > $O/r2_trace_k512_d128.txt 2>&1
tail -n 12 $O/r2_trace_k512_d64.txt
