TIERS=3 python tools/sites_probe.py 16 100 3 2>&1 | grep "sites="
TIERS=2 python tools/sites_probe.py 16 100 2 2>&1 | grep "sites="
TIERS=3 PHASES=1 python tools/sites_probe.py 16 100 3 2>&1 | grep -v "final obj"
