python tools/probe_km_dense.py 2>&1 | tail -5
