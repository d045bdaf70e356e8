#!/bin/bash
for cfg in 230 231 220 221 140 141 130 131 160 161; do A3D_FUSED_ADAM_CFG=$cfg timeout 120 python tools/adam_probe.py 2>&1 | tail -1 | cut -c1-250; done
