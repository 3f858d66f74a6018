for cfg in 0 1 2 3 4 5 6 7; do echo CFG=$cfg; ACR_PAMR_CFG=$cfg python scripts/bench_refine.py pamr 2>&1 | tail -1 | cut -c1-200; done
