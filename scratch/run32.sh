timeout 600 python scratch/occ_check.py 2>&1 | tail -40
