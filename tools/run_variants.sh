# development helper: long single-pair calls against the fused item cap
for c in 1048576 262144 131072 65536 32768; do echo "== SATMC_FUSED_MAX_CHUNK=$c"; SATMC_FUSED_MAX_CHUNK=$c python tools/probe_long.py 12.5e9 25e9 2e9 2>&1 | grep -E "pair x|rror"; done
