for d in 5 6; do echo DBG=$d; ACR_DBG=$d timeout 100 python scripts/bench_attn.py 2>&1 | tail -1 | grep -o '"bwd_noG_us": [0-9.]*\|"bwd_codes_us": [0-9.]*' | tr '\n' ' '; echo; done
