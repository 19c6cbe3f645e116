python scratch/c5_probe.py 2>&1 | tail -12
