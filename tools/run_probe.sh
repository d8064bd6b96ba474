python tools/sites_var.py 2>&1 | grep rep
