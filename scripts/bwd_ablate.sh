for d in 0 1 2 4 3 7; do echo DBG=$d; ACR_BWD_DBG=$d python scripts/bench_attn.py 2>&1 | tail -1 | grep -o '"bwd_noG_us": [0-9.]*'; done
