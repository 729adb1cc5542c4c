# development helper: small-call latency, planner's own choice (0) against the resident cut (2)
for b in 0 2; do echo "== SATMC_TINY_BPS=$b"; SATMC_TINY_BPS=$b python tools/probe_cfg2.py 2>&1 | grep -E "pair x|rror"; done
