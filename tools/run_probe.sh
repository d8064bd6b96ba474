PHASES=1 PARTS=6 python tools/sites_probe.py 16 100 6 2>&1 | grep -v "final obj"
