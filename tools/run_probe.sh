CPU=1 python tools/configs_probe.py 2>&1 | tail -3
