SHARED=1 python tools/sites_probe.py 16 100 4 2>&1 | grep "sites="
SHARED=0 python tools/sites_probe.py 16 100 4 2>&1 | grep "sites="
SHARED=1 PHASES=1 python tools/sites_probe.py 16 100 4 2>&1 | grep -v "final obj"
