python tools/probe_concurrent_chunks.py 2>&1 | tail -8
