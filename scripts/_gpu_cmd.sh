timeout 300 python scripts/exp_batch_signs.py 2>&1 | tail -5
