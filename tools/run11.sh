#!/bin/bash
echo "== 1 CTA / SM (DGP_SMEM_PAD=20000)"
DGP_SMEM_PAD=20000 python tools/tile_probe.py 2>&1
