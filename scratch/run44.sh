python scratch/host_profile.py 2>&1 | tail -45
