python tools/probe_policy.py 2>&1 | tail -3
