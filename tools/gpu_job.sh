#!/bin/bash
mkdir -p gpurun_out
python tools/percal_probe.py inception_v3 50 2>&1 | grep "B=32" 
SPAA_GRAPH_POOL=0 python tools/percal_probe.py inception_v3 50 2>&1 | grep "B=32"
